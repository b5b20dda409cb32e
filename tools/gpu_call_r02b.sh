#!/bin/bash
# Round 2, second GPU call: the tiled solve kernel.  Full -m gpu suite, then the headline bench with the tiled kernel and
# with round 1's owned kernel side by side, then the 100k mesh.
O=gpurun_out
mkdir -p $O
rm -f $O/parity_report.jsonl
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/r02b_pytest.log 2>&1; echo "pytest rc $?" >> $O/r02b_pytest.log
cp $O/parity_report.jsonl $O/r02b_parity.jsonl 2>/dev/null
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02b_bench_1m_tiled.json 2> $O/r02b_bench_1m_tiled.err
ADMM_B200_GS_TILED=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02b_bench_1m_owned.json 2> $O/r02b_bench_1m_owned.err
timeout 600 python bench.py --steps 10 --warmup 3 --workload beam_100k --no-cpu-baseline > $O/r02b_bench_100k_tiled.json 2> $O/r02b_bench_100k_tiled.err
tail -5 $O/r02b_pytest.log
for f in $O/r02b_bench_*.json; do echo "== $f"; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['step_breakdown_ms'], {k:round(v['ms_per_launch'],4) for k,v in d['kernels'].items()}, d['details']['global_solve_kernel'][-70:])
"; done
