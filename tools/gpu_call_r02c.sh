#!/bin/bash
O=gpurun_out
for wl in beam_1m beam_100k; do for dbg in 0 2 16 18; do timeout 300 python tools/gs_prof2.py $wl $dbg 2>&1 | grep -v "^$"; done; done > $O/r02c_gsprof2.log 2>&1
cat $O/r02c_gsprof2.log
