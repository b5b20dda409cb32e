#!/bin/bash
# 2 GPUs: the multi-rank parity test (log kept) and the bench with its built-in comparison against a single-GPU run
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/r02h_smi.txt
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -rs -s -p no:cacheprovider > $O/r02h_pytest_n2.log 2>&1; echo "pytest rc $?" >> $O/r02h_pytest_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r02h_bench_1m_n2.json 2> $O/r02h_bench_1m_n2.err; echo "bench rc $?" >> $O/r02h_bench_1m_n2.err
tail -4 $O/r02h_pytest_n2.log
tail -3 $O/r02h_bench_1m_n2.err
python -c "
import json
d=json.loads([l for l in open('$O/r02h_bench_1m_n2.json').read().splitlines() if l.startswith('{')][-1])
print(d['value'], d['e2e']['value'], d.get('parity_vs_n1'), d['step_breakdown_ms'], d['details']['n_elements_this_rank'])
"
