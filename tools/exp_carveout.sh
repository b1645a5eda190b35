run() { env $1 timeout 300 python bench.py --steps 20 --warmup 10 --no-cpu-baseline --no-ppo --arenas $2 > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
python -c "
import json; b=json.load(open('gpurun_out/ab.json')); print('[$1 arenas=$2]', 'value %.3fM' % (b['value']/1e6), 'k_roles %.3f ms' % b['roofline']['launch_ms'], 'us/arena-launch %.4f' % (1e3*b['roofline']['launch_ms']/$2))"; }
run A=1 16384
run A=1 14208
run RLG_SMEM_CARVEOUT=71 14208
run A=1 9472
run RLG_SMEM_CARVEOUT=57 9472
run A=1 4736
run RLG_SMEM_CARVEOUT=28 4736
run A=1 32768
run A=1 65536
