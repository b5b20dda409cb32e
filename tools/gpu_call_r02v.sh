#!/bin/bash
O=gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02v_bench_1m.json 2> $O/r02v_bench_1m.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload beam_100k > $O/r02v_bench_100k.json 2> $O/r02v_bench_100k.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --model 2 --floor > $O/r02v_bench_c3.json 2> $O/r02v_bench_c3.err
for f in $O/r02v_bench_*.json; do echo "== $f"; python -c "
import json,sys
t=open('$f').read().strip()
if not t: print('EMPTY'); sys.exit()
d=json.loads(t.splitlines()[-1])
print(round(d['value'],1), round(d['e2e']['value'],1), {k:round(v,3) for k,v in d['step_breakdown_ms'].items()}, {k:(round(v['ms_per_launch'],4), round(v['frac'],3)) for k,v in d['kernels'].items()})
"; done
