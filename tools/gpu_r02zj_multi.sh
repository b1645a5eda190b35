#!/bin/bash
# r02zj (8-GPU box): final weak-scaling line at 8 GPUs and the strong-scaling line of the 16 384-arena pool on the final code
mkdir -p gpurun_out
run() {
  n=$1; tag=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) bench.py --gpus $n "$@" > gpurun_out/r02zj_$tag.json 2> gpurun_out/r02zj_$tag.err
  echo "$tag rc=$? $(python - <<PY
import json
try:
    b=json.loads(open('gpurun_out/r02zj_$tag.json').read().strip().splitlines()[-1])
    print('value %.2fM' % (b['value']/1e6), 'ms/step %.3f' % b['ms_per_step'], 'e2e %.2fM' % (b['e2e']['value']/1e6), 'k_roles %.3f' % b['roofline']['launch_ms'], 'ppo', (b.get('ppo_iteration') or {}).get('total_iteration_time_s'))
except Exception as ex: print('no line', ex)
PY
)" | tee -a gpurun_out/r02zj_summary.txt
}
rm -f gpurun_out/r02zj_summary.txt
run 8 weak8_cfg2 --steps 200 --warmup 40 --no-cpu-baseline
run 8 strong16k_8 --total-arenas 16384 --steps 150 --warmup 40 --no-cpu-baseline --no-ppo
run 4 strong16k_4 --total-arenas 16384 --steps 150 --warmup 40 --no-cpu-baseline --no-ppo
run 2 strong16k_2 --total-arenas 16384 --steps 150 --warmup 40 --no-cpu-baseline --no-ppo
