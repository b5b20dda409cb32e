#!/bin/bash
O=gpurun_out
timeout 300 python tools/ldlt_prof.py cloth_512 > $O/r02m_ldlt_prof_cloth.txt 2>&1
timeout 300 python tools/ldlt_prof.py beam_100k > $O/r02m_ldlt_prof_beam100k.txt 2>&1
cat $O/r02m_ldlt_prof_cloth.txt $O/r02m_ldlt_prof_beam100k.txt
