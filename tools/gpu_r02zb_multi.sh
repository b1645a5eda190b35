#!/bin/bash
# r02zb (8-GPU box): the 2-rank NCCL test of the device PPO update, and BASELINE configs[1] weak scaling at 8 GPUs on the final code
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ppo.py -m gpu -q -k "two_ranks" > gpurun_out/pytest_r02zb.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r02zb.log
tail -3 gpurun_out/pytest_r02zb.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 200 --warmup 40 > gpurun_out/r02zb_weak8_cfg2.json 2> gpurun_out/r02zb_weak8.err; echo "bench rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 8 --impl reference --steps 3 --warmup 3 > gpurun_out/r02zb_weak8_reference.json 2> gpurun_out/r02zb_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
b=json.loads(open('gpurun_out/r02zb_weak8_cfg2.json').read().strip().splitlines()[-1])
print('8 GPUs: value %.2fM' % (b['value']/1e6), 'e2e %.2fM' % (b['e2e']['value']/1e6), 'ppo', (b.get('ppo_iteration') or {}).get('total_iteration_time_s'), 'cpu', b.get('cpu_baseline'))
try:
    r=json.loads(open('gpurun_out/r02zb_weak8_reference.json').read().strip().splitlines()[-1]); print('reference arm %.3fM on %s cores' % (r['value']/1e6, r['cpu_baseline']['cores']))
except Exception as ex: print('ref', ex)
PY
