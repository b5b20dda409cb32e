#!/bin/bash
# Round 2 profiling call: GPU tests (incl. bunny), launch lists and ncu --set full captures of every dominant kernel.
O=gpurun_out
mkdir -p $O
rm -f $O/parity_report.jsonl
# (the -m gpu suite ran in the first attempt of this call: 129 passed, 3 skipped for lack of a second GPU)
# every .ncu-rep is reduced to its summaries on the box: seven full reports exceed what gpurun copies back
sumrep() { bash profiles/ncusum.sh $O/$1.ncu-rep > $O/$1_summary.txt 2>&1; ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1_raw.csv 2>/dev/null; [ "$2" = keep ] || rm -f $O/$1.ncu-rep; }
# launch list of the default bench command (headline), serialised / cold: the SHARES must agree with the CUDA-event breakdown
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file $O/r02j_launches_1m.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r02j_ncu_bench_1m.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mcgs_owned -s 20 -c 1 -o $O/r02j_mcgs_owned -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
sumrep r02j_mcgs_owned keep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tet_local_kernel -s 20 -c 1 -o $O/r02j_tet_local -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
sumrep r02j_tet_local drop
timeout 900 ncu --set full --clock-control none --import-source on -k regex:assemble_kernel -s 20 -c 1 -o $O/r02j_assemble -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
sumrep r02j_assemble drop
# C2 (LDLT, 100k) and C4 (cloth): launch list + the block solve and the triangle kernel
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 300 --csv --log-file $O/r02j_launches_c2.csv python bench.py --steps 2 --warmup 3 --workload beam_100k --linsolver 0 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ldlt_blocks -s 20 -c 1 -o $O/r02j_ldlt_blocks_c2 -f python bench.py --steps 2 --warmup 3 --workload beam_100k --linsolver 0 --no-cpu-baseline > /dev/null 2>&1
sumrep r02j_ldlt_blocks_c2 drop
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 300 --csv --log-file $O/r02j_launches_c4.csv python bench.py --steps 2 --warmup 3 --workload cloth_512 --no-cpu-baseline > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ldlt_blocks -s 20 -c 1 -o $O/r02j_ldlt_blocks_c4 -f python bench.py --steps 2 --warmup 3 --workload cloth_512 --no-cpu-baseline > /dev/null 2>&1
sumrep r02j_ldlt_blocks_c4 drop
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tri_local_kernel -s 20 -c 1 -o $O/r02j_tri_local -f python bench.py --steps 2 --warmup 3 --workload cloth_512 --no-cpu-baseline > /dev/null 2>&1
sumrep r02j_tri_local drop
# a clean (unprofiled) headline line with the new assemble kernel
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02j_bench_1m.json 2> $O/r02j_bench_1m.err
ls -la $O/r02j_*
python -c "
import json
d=json.loads(open('$O/r02j_bench_1m.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['step_breakdown_ms'], {k:round(v['ms_per_launch'],4) for k,v in d['kernels'].items()})
"
