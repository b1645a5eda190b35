#!/bin/bash
# round-end evidence run: bench (+ cpu baseline, ppo iteration), reference arm, launch list, full ncu captures, rollout statistics
tag=${1:-cur}
mkdir -p gpurun_out
timeout 900 python bench.py --steps 50 --warmup 10 > gpurun_out/bench_$tag.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_reference_$tag.json 2> gpurun_out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ppo > gpurun_out/ncu_bench.log 2>&1
# DRAM bytes of k_roles alone (no L2 flush: the flush buffer's write-back would be attributed to the kernel) and with the bench's flush
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_roles -s 60 -c 8 --csv --log-file gpurun_out/traffic_noflush_$tag.csv python bench.py --steps 10 --warmup 10 --no-cpu-baseline --no-ppo --no-l2-flush > /dev/null 2>&1
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_roles -s 60 -c 8 --csv --log-file gpurun_out/traffic_flush_$tag.csv python bench.py --steps 10 --warmup 10 --no-cpu-baseline --no-ppo > /dev/null 2>&1
timeout 800 ncu --set full --clock-control none --import-source on -k regex:k_roles -s 70 -c 1 -o gpurun_out/prof_k_roles_$tag -f python bench.py --steps 20 --warmup 10 --no-cpu-baseline --no-ppo > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mlp_infer -s 40 -c 1 -o gpurun_out/prof_k_mlp_$tag -f python bench.py --steps 8 --warmup 5 --no-cpu-baseline --no-ppo > gpurun_out/ncu_full_mlp.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tma -s 5 -c 1 -o gpurun_out/prof_k_gemm_$tag -f python tools/gemm_bench.py > gpurun_out/ncu_gemm.log 2>&1
timeout 300 python tools/gemm_bench.py > gpurun_out/gemm_bench_$tag.txt 2>&1
RLG_STATS_OUT=$PWD/gpurun_out/rollout_stats_$tag.json timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k rollout 2>&1 | tail -2
python -c "
import json; b=json.load(open('gpurun_out/bench_$tag.json')); print('value %.3fM' % (b['value']/1e6), 'e2e %.3fM' % (b['e2e']['value']/1e6), 'k_roles %.3f ms' % b['roofline']['launch_ms'], 'mlp %.3f ms' % b['roofline_mlp']['launch_ms'], 'cpu', b.get('cpu_baseline'), b.get('ppo_iteration'))"
ls -la gpurun_out | tail -12
