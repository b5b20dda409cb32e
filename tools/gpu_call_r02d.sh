#!/bin/bash
O=gpurun_out
mkdir -p $O
rm -f $O/parity_report.jsonl
timeout 60 tools/micro/_build/gather_bench > $O/r02d_gather_bench.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > $O/r02d_pytest.log 2>&1; echo "pytest rc $?" >> $O/r02d_pytest.log
cp $O/parity_report.jsonl $O/r02d_parity.jsonl 2>/dev/null
timeout 600 python bench.py --steps 10 --warmup 3 --workload beam_100k --linsolver 0 > $O/r02d_bench_c2_100k_ldlt.json 2> $O/r02d_bench_c2.err
timeout 900 python bench.py --steps 10 --warmup 3 --workload cloth_512 > $O/r02d_bench_c4_cloth.json 2> $O/r02d_bench_c4.err
timeout 900 python bench.py --steps 10 --warmup 3 --workload cloth_512 --limits --no-cpu-baseline > $O/r02d_bench_c4_cloth_limits.json 2> $O/r02d_bench_c4_limits.err
tail -5 $O/r02d_pytest.log
grep -v "syncthreads" $O/r02d_gather_bench.txt | grep "warps 16\|warps  1"
for f in $O/r02d_bench_*.json; do echo "== $f"; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['step_breakdown_ms'], {k:round(v['ms_per_launch'],4) for k,v in d['kernels'].items()}, d['details']['global_solve_kernel'], d.get('cpu_baseline',{}).get('value'))
"; tail -2 ${f%.json}.err 2>/dev/null; done
