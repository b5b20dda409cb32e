#!/bin/bash
# 16-byte mailbox words (x, y, z, tag in one store / one polling load): parity, then bench with and without helpers.
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_scale.py tests/test_gpu_parity.py -m gpu -x -q -k "100k or unstructured or bunny or steps or floor or obstacle or pins or beam" > $O/r02u_pytest.log 2>&1
tail -3 $O/r02u_pytest.log
for h in 1 4; do
  ADMM_B200_GS_HELP=$h timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02u_bench_1m_help$h.json 2> $O/r02u_bench_1m_help$h.err
  ADMM_B200_GS_HELP=$h timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload beam_100k > $O/r02u_bench_100k_help$h.json 2> $O/r02u_bench_100k_help$h.err
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --model 2 --floor > $O/r02u_bench_c3.json 2> $O/r02u_bench_c3.err
for f in $O/r02u_bench_*.json; do echo "== $f"; python -c "
import json,sys
t=open('$f').read().strip()
if not t: print('EMPTY'); sys.exit()
d=json.loads(t.splitlines()[-1])
print(round(d['value'],1), round(d['e2e']['value'],1), {k:round(v,3) for k,v in d['step_breakdown_ms'].items()}, {k:(round(v['ms_per_launch'],4), round(v['frac'],3)) for k,v in d['kernels'].items()})
"; done
rm -f $O/r02u_gsprof.log
for wl in beam_100k beam_1m; do
  echo "=== $wl" >> $O/r02u_gsprof.log
  ADMM_B200_GS_DBG=$((70*256)) timeout 300 python tools/gs_prof.py $wl >> $O/r02u_gsprof.log 2>&1
done
grep -n "===\|cycles per pass" $O/r02u_gsprof.log | cut -c1-330
