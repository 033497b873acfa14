# Run under `gpurun --gpus 8`: bench.py at 2 / 4 / 8 ranks for cfg3 and cfg5s; profiles/r2_scaling.json is built from the lines.
for N in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_scale_cfg3_n$N.json 2> gpurun_out/r2_scale_cfg3_n$N.err
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 5 --warmup 3 --workload cfg5s --no-svi --no-calnf > gpurun_out/r2_scale_cfg5s_n$N.json 2> gpurun_out/r2_scale_cfg5s_n$N.err
done
python - <<'PY'
import json
for w in ("cfg3","cfg5s"):
    for n in (2,4,8):
        try:
            d=json.loads(open(f"gpurun_out/r2_scale_{w}_n{n}.json").read().strip().splitlines()[-1])
            print(w,n,"value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"strong",round(d["strong"]["value"]),"svi",d.get("svi_step_ms") and {k:round(v,3) for k,v in d["svi_step_ms"].items() if isinstance(v,float)},"calnf",d.get("calnf_step_ms") and round(d["calnf_step_ms"]["value"],1))
        except Exception as e:
            print(w,n,"failed",e)
PY
