#!/bin/bash
# Round 2, first GPU call: the whole -m gpu suite (with the new production-scale tests), the gather micro-benchmark and a
# first line for every BASELINE configuration.  Everything lands in gpurun_out/r02a_*.
O=gpurun_out
mkdir -p $O
rm -f $O/parity_report.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/r02a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > $O/r02a_pytest.log 2>&1; echo "pytest rc $?" >> $O/r02a_pytest.log
cp $O/parity_report.jsonl $O/r02a_parity.jsonl 2>/dev/null
timeout 120 tools/micro/_build/gather_bench > $O/r02a_gather_bench.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02a_smoke.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > $O/r02a_bench_1m.json 2> $O/r02a_bench_1m.err
timeout 600 python bench.py --steps 10 --warmup 3 --coloring random --no-cpu-baseline > $O/r02a_bench_1m_random.json 2> $O/r02a_bench_1m_random.err
timeout 600 python bench.py --steps 10 --warmup 3 --workload beam_100k --linsolver 0 > $O/r02a_bench_c2_100k_ldlt.json 2> $O/r02a_bench_c2.err
timeout 600 python bench.py --steps 10 --warmup 3 --workload beam_100k > $O/r02a_bench_100k_mcgs.json 2> $O/r02a_bench_100k_mcgs.err
timeout 600 python bench.py --steps 10 --warmup 3 --model 2 --floor > $O/r02a_bench_c3_stvk_floor.json 2> $O/r02a_bench_c3.err
timeout 900 python bench.py --steps 5 --warmup 3 --workload cloth_512 > $O/r02a_bench_c4_cloth.json 2> $O/r02a_bench_c4.err
timeout 900 python bench.py --steps 5 --warmup 3 --workload cloth_512 --limits --no-cpu-baseline > $O/r02a_bench_c4_cloth_limits.json 2> $O/r02a_bench_c4_limits.err
tail -3 $O/r02a_pytest.log
for f in $O/r02a_bench_*.json; do echo "== $f"; head -c 400 $f; echo; done
