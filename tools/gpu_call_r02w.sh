#!/bin/bash
# 2 GPUs after the mailbox format change: sharded parity tests and the N=2 bench line with its built-in comparison.
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -rs -s -p no:cacheprovider > $O/r02w_pytest_mgpu_n2.log 2>&1; echo "pytest rc $?" >> $O/r02w_pytest_mgpu_n2.log
tail -3 $O/r02w_pytest_mgpu_n2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r02w_bench_1m_n2.json 2> $O/r02w_bench_1m_n2.err; echo "bench n2 rc $?" >> $O/r02w_bench_1m_n2.err
tail -2 $O/r02w_bench_1m_n2.err; python -c "
import json
d=json.loads([l for l in open('$O/r02w_bench_1m_n2.json').read().splitlines() if l.startswith('{')][-1])
print('N=2', round(d['value'],1), round(d['e2e']['value'],1), d.get('parity_vs_n1',{}).get('max_over_bbox'), d['step_breakdown_ms'], d['kernels']['mcgs_kernel']['ms_per_launch'])
"
