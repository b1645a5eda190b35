#!/bin/bash
# r02p (8-GPU box): BASELINE configs[1] weak scaling at 8, configs[2] (2v2 padded + zero-sum, 8 192 arenas/GPU) and configs[3] (--cfg4) at 8,
# strong scaling of fixed pools (16 384 and 65 536 arenas) over 2/4/8 GPUs.  One JSON line per run into gpurun_out/r02p_*.json
mkdir -p gpurun_out
run() {  # run N tag args...
  n=$1; tag=$2; shift 2
  if [ "$n" = 1 ]; then timeout 600 python bench.py --gpus 1 "$@" > gpurun_out/r02p_$tag.json 2> gpurun_out/r02p_$tag.err
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) bench.py --gpus $n "$@" > gpurun_out/r02p_$tag.json 2> gpurun_out/r02p_$tag.err; fi
  echo "$tag rc=$? $(python - <<PY
import json
try:
    b=json.loads(open('gpurun_out/r02p_$tag.json').read().strip().splitlines()[-1])
    print('value %.2fM' % (b['value']/1e6), 'ms/step %.3f' % b['ms_per_step'], 'e2e', ('%.2fM' % (b['e2e']['value']/1e6)) if 'e2e' in b else '-', 'ppo', (b.get('ppo_iteration') or {}).get('total_iteration_time_s'), b.get('iteration'))
except Exception as ex: print('no line', ex)
PY
)" | tee -a gpurun_out/r02p_summary.txt
}
rm -f gpurun_out/r02p_summary.txt
run 8 weak8_cfg2 --steps 200 --warmup 40 --no-cpu-baseline
run 8 cfg3_8gpu --team 2 --padded-obs --zero-sum --arenas 8192 --steps 200 --warmup 40 --no-cpu-baseline
run 8 cfg4_8gpu --cfg4 --arenas 8192 --steps 12 --warmup 4
for n in 2 4 8; do
  run $n strong16k_$n --total-arenas 16384 --steps 150 --warmup 40 --no-cpu-baseline --no-ppo
  run $n strong64k_$n --total-arenas 65536 --steps 100 --warmup 40 --no-cpu-baseline --no-ppo
done
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv | head -9
