#!/bin/bash
O=gpurun_out
mkdir -p $O
rm -f $O/parity_report.jsonl
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > $O/r02l_pytest.log 2>&1; echo "pytest rc $?" >> $O/r02l_pytest.log
tail -5 $O/r02l_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --workload beam_100k --linsolver 0 --no-cpu-baseline > $O/r02l_bench_c2_100k_ldlt.json 2> $O/r02l_bench_c2.err
timeout 900 python bench.py --steps 10 --warmup 3 --workload cloth_512 --no-cpu-baseline > $O/r02l_bench_c4_cloth.json 2> $O/r02l_bench_c4.err
for f in $O/r02l_bench_*.json; do echo "== $f"; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['step_breakdown_ms'], {k:round(v['ms_per_launch'],4) for k,v in d['kernels'].items()}, d['details']['global_solve_kernel'][-160:])
"; tail -2 ${f%.json}.err 2>/dev/null; done
