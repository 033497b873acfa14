python -m pytest tests/test_cuda_parity.py -m gpu -x -q -k "large_networks or golden or oracle" > gpurun_out/r2_pytest14.log 2>&1; tail -2 gpurun_out/r2_pytest14.log
python tools/quick_bench.py cfg5s 512 0.01 3 2>&1 | tail -1
python tools/quick_bench.py cfg5s 1332 0.01 3 2>&1 | tail -1
python tools/quick_bench.py cfg3 4096 0.012 3 2>&1 | tail -1
python tools/quick_bench.py cfg2 1024 0.012 5 2>&1 | tail -1
true 2>&1 | tail -1
