#!/bin/bash
set -x
mkdir -p gpurun_out/t
for c in 12 16 20 24 28 32 48; do RT_B200_SPARSE_CTAS=$c timeout 300 python tools/kstep_ab.py --label "sparse_ctas=$c" >> gpurun_out/t/sparse_ctas_sweep.txt 2>&1; done
for h in 19 3; do RT_B200_SPARSE_HINTS=$h timeout 300 python tools/kstep_ab.py --label "sparse_hints=$h (bit4: plain st.global instead of .cs)" >> gpurun_out/t/sparse_ctas_sweep.txt 2>&1; done
cat gpurun_out/t/sparse_ctas_sweep.txt
