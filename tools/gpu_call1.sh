python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest2.log 2>&1; tail -5 gpurun_out/r2_pytest2.log
python bench.py > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; cat gpurun_out/r2_bench2.json; tail -3 gpurun_out/r2_bench2.err
BA_PHASE_PROFILE=1 python tools/profile_forward.py 4096 0.012 10 2 2>&1 | tail -4
BA_PHASE_PROFILE=1 python tools/profile_forward.py 4096 0.05 10 2 2>&1 | tail -3
bash tools/make_profiles_r2.sh
