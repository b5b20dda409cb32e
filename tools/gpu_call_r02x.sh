#!/bin/bash
# Block LDLT solve with the first descriptor of every phase asked for before the barrier, and the row's own value before the dot product.
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py tests/test_gpu_binding.py -m gpu -x -q -k "ldlt or linsolve or uzawa or cloth or golden or bunny or unstructured or reference_step or lineartet" > $O/r02x_pytest.log 2>&1
tail -3 $O/r02x_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload beam_100k --linsolver 0 > $O/r02x_bench_c2.json 2> $O/r02x_bench_c2.err
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload cloth_512 > $O/r02x_bench_c4.json 2> $O/r02x_bench_c4.err
for f in $O/r02x_bench_*.json; do echo "== $f"; python -c "
import json,sys
t=open('$f').read().strip()
if not t: print('EMPTY'); sys.exit()
d=json.loads(t.splitlines()[-1])
print(round(d['value'],1), round(d['e2e']['value'],1), {k:round(v,3) for k,v in d['step_breakdown_ms'].items()}, {k:(round(v['ms_per_launch'],4), round(v['frac'],3)) for k,v in d['kernels'].items()})
"; done
timeout 300 python tools/ldlt_prof.py beam_100k > $O/r02x_ldlt_prof_beam100k.txt 2>&1; tail -8 $O/r02x_ldlt_prof_beam100k.txt
timeout 300 python tools/ldlt_prof.py cloth_512 > $O/r02x_ldlt_prof_cloth.txt 2>&1; tail -4 $O/r02x_ldlt_prof_cloth.txt
