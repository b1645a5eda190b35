#!/bin/bash
# r02zq (8-GPU box): configs[2] and configs[3] at 8 GPUs on the final code
mkdir -p gpurun_out
rm -f gpurun_out/r02zq_summary.txt
run() {
  n=$1; tag=$2; shift 2
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) bench.py --gpus $n "$@" > gpurun_out/r02zq_$tag.json 2> gpurun_out/r02zq_$tag.err
  echo "$tag rc=$? $(python - <<PY
import json
try:
    b=json.loads(open('gpurun_out/r02zq_$tag.json').read().strip().splitlines()[-1])
    print('value %.2fM' % (b['value']/1e6), 'ms/step %.3f' % b['ms_per_step'], 'e2e', ('%.2fM' % (b['e2e']['value']/1e6)) if 'e2e' in b else '-', 'ppo', (b.get('ppo_iteration') or {}).get('total_iteration_time_s'), b.get('iteration'))
except Exception as ex: print('no line', ex)
PY
)" | tee -a gpurun_out/r02zq_summary.txt
}
run 8 cfg3_8gpu --team 2 --padded-obs --zero-sum --arenas 8192 --steps 150 --warmup 40 --no-cpu-baseline
run 8 cfg4_8gpu --cfg4 --arenas 8192 --steps 12 --warmup 4
