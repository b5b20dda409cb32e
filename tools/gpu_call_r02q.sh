#!/bin/bash
# Helpers in the static-ownership solve kernel (ADMM_B200_GS_HELP = ways a boundary slice's rows are split): parity first, then A/B.
O=gpurun_out
mkdir -p $O
ADMM_B200_GS_HELP=4 timeout 900 python -m pytest tests/test_gpu_scale.py tests/test_gpu_parity.py -m gpu -x -q -k "100k or unstructured or bunny or steps or floor or obstacle or pins" > $O/r02q_pytest_help4.log 2>&1
tail -5 $O/r02q_pytest_help4.log
for h in 1 2 3 4; do
  ADMM_B200_GS_HELP=$h timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02q_bench_1m_help$h.json 2> $O/r02q_bench_1m_help$h.err
done
for h in 1 2 4; do
  ADMM_B200_GS_HELP=$h timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload beam_100k > $O/r02q_bench_100k_help$h.json 2> $O/r02q_bench_100k_help$h.err
done
ADMM_B200_GS_HELP=4 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --model 2 --floor > $O/r02q_bench_c3_help4.json 2> $O/r02q_bench_c3_help4.err
for f in $O/r02q_bench_*.json; do echo "== $f"; python -c "
import json,sys
t=open('$f').read().strip()
if not t: print('EMPTY'); sys.exit()
d=json.loads(t.splitlines()[-1])
print(round(d['value'],1), round(d['e2e']['value'],1), {k:round(v,3) for k,v in d['step_breakdown_ms'].items()}, {k:(round(v['ms_per_launch'],4), round(v['frac'],3)) for k,v in d['kernels'].items()})
"; done
tail -3 $O/r02q_bench_1m_help4.err
