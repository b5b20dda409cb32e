#!/bin/bash
O=gpurun_out
rm -f $O/r02s_gsprof.log
for wl in beam_100k beam_1m; do for h in 1; do
  echo "=== $wl help $h" >> $O/r02s_gsprof.log
  ADMM_B200_GS_HELP=$h ADMM_B200_GS_DBG=$((70*256)) timeout 300 python tools/gs_prof.py $wl >> $O/r02s_gsprof.log 2>&1
done; done
