# strong / weak scaling of cfg3 only (no SVI / CalNF / predictive sections): 8 GPUs
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 5 --warmup 3 --no-svi --no-calnf --no-predictive > gpurun_out/r2_scale_quick_n8.json 2> gpurun_out/r2_scale_quick_n8.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_scale_quick_n8.json').read().strip().splitlines()[-1])
print('n8 value', d['value'], 'strong', d['strong'])"
