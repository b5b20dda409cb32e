#!/bin/bash
O=gpurun_out
for dbg in 0 128 256 384; do
ADMM_B200_GS_DBG=$dbg timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02f_bench_1m_dbg$dbg.json 2> $O/r02f_bench_1m_dbg$dbg.err
python -c "
import json,sys
d=json.loads(open('$O/r02f_bench_1m_dbg$dbg.json').read().strip().splitlines()[-1])
print('dbg $dbg', d['value'], {k:round(v['ms_per_launch'],4) for k,v in d['kernels'].items()})
"
done
