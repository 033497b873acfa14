python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest3.log 2>&1; tail -15 gpurun_out/r2_pytest3.log
python bench.py --steps 5 > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err; cat gpurun_out/r2_bench3.json; tail -5 gpurun_out/r2_bench3.err
