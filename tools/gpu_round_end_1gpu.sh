# Run under `gpurun`: GPU test suite, bench lines (both arms, cfg3 and cfg5s), phase and occupancy tables, ncu captures, smoke.
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_final.log 2>&1; tail -3 gpurun_out/r2_pytest_final.log
python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; tail -2 gpurun_out/r2_bench_final.err; cat gpurun_out/r2_bench_final.json | cut -c1-400
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_final_ref.json 2>/dev/null; cut -c1-200 gpurun_out/r2_bench_final_ref.json
python bench.py --workload cfg5s --steps 5 --no-svi --no-calnf > gpurun_out/r2_bench_cfg5s.json 2>/dev/null; cut -c1-200 gpurun_out/r2_bench_cfg5s.json
python tools/quick_bench.py cfg2 1024 0.012 5 2>&1 | tail -1
python tools/quick_bench.py cfg1 1024 0.03 5 2>&1 | tail -1
BA_PHASE_PROFILE=1 python tools/profile_forward.py 4096 0.012 10 2 2>&1 | tail -2
for g in 1184 2368 3552; do BA_GRID_CAP=$g python tools/quick_bench.py cfg3 4096 0.012 3 2>&1 | tail -1; done
bash tools/make_profiles_r2.sh
python -c "import __graft_entry__ as g; g.smoke()"
