python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest4.log 2>&1; tail -4 gpurun_out/r2_pytest4.log
BA_PHASE_PROFILE=1 python tools/profile_forward.py 4096 0.012 10 3 2>&1 | tail -3
python tools/profile_forward.py 4096 0.012 10 3 2>&1 | tail -1
python tools/profile_forward.py 4096 0.05 10 3 2>&1 | tail -1
