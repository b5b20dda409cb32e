#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "not scale and not prox" > $O/r02n_pytest.log 2>&1; echo "pytest rc $?" >> $O/r02n_pytest.log
tail -4 $O/r02n_pytest.log
timeout 300 python tools/ldlt_prof.py cloth_512 > $O/r02n_ldlt_prof_cloth.txt 2>&1
timeout 300 python tools/ldlt_prof.py beam_100k > $O/r02n_ldlt_prof_beam100k.txt 2>&1
grep "total\|solve ms\|forward :\|backward:" $O/r02n_ldlt_prof_cloth.txt $O/r02n_ldlt_prof_beam100k.txt
