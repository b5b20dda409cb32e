#!/bin/bash
# Final single-GPU evidence of round 2: the whole -m gpu suite, smoke(), the bench lines of every BASELINE configuration with
# their CPU baselines, launch lists and ncu --set full captures of the two solve kernels that changed since r02j.
O=gpurun_out
mkdir -p $O
rm -f $O/parity_report.jsonl
timeout 1500 python -m pytest tests -m gpu -q -rs -p no:cacheprovider > $O/r03a_pytest.log 2>&1; echo "pytest rc $?" >> $O/r03a_pytest.log
tail -4 $O/r03a_pytest.log
cp $O/parity_report.jsonl $O/r03a_parity.jsonl 2>/dev/null
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r03a_smoke.log 2>&1; tail -2 $O/r03a_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r03a_bench_1m.json 2> $O/r03a_bench_1m.err
timeout 600 python bench.py --steps 10 --warmup 3 --coloring random --no-cpu-baseline > $O/r03a_bench_1m_random.json 2> $O/r03a_bench_1m_random.err
timeout 600 python bench.py --steps 10 --warmup 3 --workload beam_100k --linsolver 0 > $O/r03a_bench_c2_100k_ldlt.json 2> $O/r03a_bench_c2.err
timeout 600 python bench.py --steps 10 --warmup 3 --workload beam_100k > $O/r03a_bench_100k_mcgs.json 2> $O/r03a_bench_100k_mcgs.err
timeout 600 python bench.py --steps 10 --warmup 3 --model 2 --floor > $O/r03a_bench_c3_stvk_floor.json 2> $O/r03a_bench_c3.err
timeout 900 python bench.py --steps 10 --warmup 3 --workload cloth_512 > $O/r03a_bench_c4_cloth.json 2> $O/r03a_bench_c4.err
sumrep() { bash profiles/ncusum.sh $O/$1.ncu-rep > $O/$1_summary.txt 2>&1; ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1_raw.csv 2>/dev/null; rm -f $O/$1.ncu-rep; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file $O/r03a_launches_1m.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mcgs_owned -s 20 -c 1 -o $O/r03a_mcgs_owned -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
sumrep r03a_mcgs_owned
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 300 --csv --log-file $O/r03a_launches_c2.csv python bench.py --steps 2 --warmup 3 --workload beam_100k --linsolver 0 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ldlt_blocks -s 20 -c 1 -o $O/r03a_ldlt_blocks_c2 -f python bench.py --steps 2 --warmup 3 --workload beam_100k --linsolver 0 --no-cpu-baseline > /dev/null 2>&1
sumrep r03a_ldlt_blocks_c2
for f in $O/r03a_bench_*.json; do echo "== $f"; python -c "
import json,sys
t=open('$f').read().strip()
if not t: print('EMPTY'); sys.exit()
d=json.loads(t.splitlines()[-1])
print(round(d['value'],1), round(d['e2e']['value'],1), {k:round(v,3) for k,v in d['step_breakdown_ms'].items()}, {k:(round(v['ms_per_launch'],4), round(v['frac'],3)) for k,v in d['kernels'].items()}, 'cpu', d.get('cpu_baseline',{}).get('value'), d.get('cpu_baseline',{}).get('cores'))
"; done
ls -la $O/r03a_* | awk '{print $5, $9}'
