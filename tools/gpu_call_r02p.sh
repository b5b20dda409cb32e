#!/bin/bash
# Round 2 final single-GPU lines: every BASELINE configuration with its CPU baseline, plus the reference arm as the driver runs it.
O=gpurun_out
mkdir -p $O
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r02p_bench_1m.json 2> $O/r02p_bench_1m.err
timeout 600 python bench.py --steps 10 --warmup 3 --coloring random --no-cpu-baseline > $O/r02p_bench_1m_random.json 2> $O/r02p_bench_1m_random.err
timeout 600 python bench.py --steps 10 --warmup 3 --workload beam_100k --linsolver 0 > $O/r02p_bench_c2_100k_ldlt.json 2> $O/r02p_bench_c2.err
timeout 600 python bench.py --steps 10 --warmup 3 --workload beam_100k > $O/r02p_bench_100k_mcgs.json 2> $O/r02p_bench_100k_mcgs.err
timeout 600 python bench.py --steps 10 --warmup 3 --model 2 --floor > $O/r02p_bench_c3_stvk_floor.json 2> $O/r02p_bench_c3.err
timeout 900 python bench.py --steps 10 --warmup 3 --workload cloth_512 > $O/r02p_bench_c4_cloth.json 2> $O/r02p_bench_c4.err
timeout 900 python bench.py --steps 10 --warmup 3 --workload cloth_512 --limits --no-cpu-baseline > $O/r02p_bench_c4_cloth_limits.json 2> $O/r02p_bench_c4_limits.err
( time timeout 1200 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r02p_bench_1m_reference.json 2> $O/r02p_bench_1m_reference.err ) 2> $O/r02p_reference_time.txt
for f in $O/r02p_bench_*.json; do echo "== $f"; python -c "
import json,sys
t=open('$f').read().strip()
if not t: print('EMPTY'); sys.exit()
d=json.loads(t.splitlines()[-1])
if d.get('impl')=='reference': print(d['value'], d['config'], d['cpu_baseline']['sample'][:160]); sys.exit()
print(round(d['value'],1), round(d['e2e']['value'],1), {k:round(v,3) for k,v in d['step_breakdown_ms'].items()}, {k:(round(v['ms_per_launch'],4), round(v['frac'],3)) for k,v in d['kernels'].items()}, 'cpu', d.get('cpu_baseline',{}).get('value'), d.get('cpu_baseline',{}).get('cores'))
"; done; cat $O/r02p_reference_time.txt
