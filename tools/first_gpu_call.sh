#!/bin/bash
# One gpurun call that answers the open questions of DESIGN.md 9.1 (about 2 GPU-minutes).  Everything under `timeout`.
#   /usr/local/graft/bin/gpurun --timeout 400 -- 'bash tools/first_gpu_call.sh'
mkdir -p gpurun_out
timeout 150 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/q_pytest.log
# publish -> halo-complete latency inside the solve kernel (global timer), plain and with a fence after the publish
for d in 0 64; do ADMM_B200_GS_DBG=$((70 * 256 + d)) timeout 100 python tools/gs_prof.py 2>&1 | grep -E "global timer|cycles per pass|kernel phases|solve us" ; done > gpurun_out/q_gsprof.log 2>&1
# the communication skeleton alone: neighbours x message size x delay x polling strategy
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/halo_ring tools/micro/halo_ring.cu && timeout 120 /tmp/halo_ring > gpurun_out/q_halo_ring.log 2>&1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pingpong tools/micro/pingpong.cu && timeout 60 /tmp/pingpong > gpurun_out/q_pingpong.log 2>&1
cat gpurun_out/q_pytest.log gpurun_out/q_gsprof.log; head -40 gpurun_out/q_halo_ring.log
