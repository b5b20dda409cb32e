#!/bin/bash
# 8 GPUs: the headline mesh sharded over 8 and over 4 ranks (with the built-in comparison against a single-GPU run and, at
# N = 8, the weak-scaling companion: 8M tets, 1M per GPU), and the 2-rank parity tests once more for the log.
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/r03d_smi.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 10 --warmup 3 > $O/r03d_bench_1m_n8.json 2> $O/r03d_bench_1m_n8.err; echo "bench n8 rc $?" >> $O/r03d_bench_1m_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 4 --steps 10 --warmup 3 > $O/r03d_bench_1m_n4.json 2> $O/r03d_bench_1m_n4.err; echo "bench n4 rc $?" >> $O/r03d_bench_1m_n4.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29563 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r03d_bench_1m_n2.json 2> $O/r03d_bench_1m_n2.err; echo "bench n2 rc $?" >> $O/r03d_bench_1m_n2.err
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -rs -s -p no:cacheprovider > $O/r03d_pytest_mgpu.log 2>&1; echo "pytest rc $?" >> $O/r03d_pytest_mgpu.log
tail -3 $O/r03d_pytest_mgpu.log
for n in 8 4 2; do tail -2 $O/r03d_bench_1m_n$n.err; python -c "
import json
d=json.loads([l for l in open('$O/r03d_bench_1m_n$n.json').read().splitlines() if l.startswith('{')][-1])
print('N=$n', round(d['value'],1), round(d['e2e']['value'],1), d['e2e']['h2d_bytes_per_step'], d.get('parity_vs_n1',{}).get('max_over_bbox'), d['step_breakdown_ms'], d.get('weak_8m'))
"; done
