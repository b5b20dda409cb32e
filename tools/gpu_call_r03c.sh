#!/bin/bash
# Final single-GPU bench lines (sampled timers) of every BASELINE configuration with their CPU baselines.
O=gpurun_out
mkdir -p $O
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r03c_bench_1m.json 2> $O/r03c_bench_1m.err
timeout 600 python bench.py --steps 12 --warmup 3 --coloring random --no-cpu-baseline > $O/r03c_bench_1m_random.json 2> $O/r03c_bench_1m_random.err
timeout 600 python bench.py --steps 20 --warmup 5 --workload beam_100k --linsolver 0 > $O/r03c_bench_c2_100k_ldlt.json 2> $O/r03c_bench_c2.err
timeout 600 python bench.py --steps 20 --warmup 5 --workload beam_100k > $O/r03c_bench_100k_mcgs.json 2> $O/r03c_bench_100k_mcgs.err
timeout 600 python bench.py --steps 20 --warmup 5 --model 2 --floor > $O/r03c_bench_c3_stvk_floor.json 2> $O/r03c_bench_c3.err
timeout 900 python bench.py --steps 20 --warmup 5 --workload cloth_512 > $O/r03c_bench_c4_cloth.json 2> $O/r03c_bench_c4.err
timeout 900 python bench.py --steps 12 --warmup 3 --workload cloth_512 --limits --no-cpu-baseline > $O/r03c_bench_c4_cloth_limits.json 2> $O/r03c_bench_c4_limits.err
python tools/timer_overhead.py > $O/r03c_timer_overhead.txt 2>&1
for f in $O/r03c_bench_*.json; do echo "== $f"; python -c "
import json,sys
t=open('$f').read().strip()
if not t: print('EMPTY'); sys.exit()
d=json.loads(t.splitlines()[-1])
print(round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['step_breakdown_ms'].items()}, {k:(round(v['ms_per_launch'],4), round(v['frac'],3)) for k,v in d['kernels'].items()}, 'cpu', d.get('cpu_baseline',{}).get('value'), d.get('cpu_baseline',{}).get('cores'), d['clocks'])
"; done; tail -4 $O/r03c_timer_overhead.txt
