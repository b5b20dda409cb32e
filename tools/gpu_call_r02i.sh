#!/bin/bash
O=gpurun_out
mkdir -p $O
rm -f $O/parity_report.jsonl
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/r02i_pytest.log 2>&1; echo "pytest rc $?" >> $O/r02i_pytest.log
cp $O/parity_report.jsonl $O/r02i_parity.jsonl 2>/dev/null
tail -30 $O/r02i_pytest.log | cut -c1-250
