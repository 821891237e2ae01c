#!/bin/bash
set -x
mkdir -p gpurun_out/u
for c in 2 4 6 8 10 12 14 24; do RT_B200_SPARSE_CTAS=$c timeout 300 python tools/kstep_ab.py --label "c3 sparse_ctas=$c" >> gpurun_out/u/sparse_ctas_sweep2.txt 2>&1; done
for c in 6 8 12 24; do RT_B200_OBS_COMPRESS=0 RT_B200_SPARSE_CTAS=$c timeout 300 python tools/kstep_ab.py --label "c3 plain obs sparse_ctas=$c" >> gpurun_out/u/sparse_ctas_sweep2.txt 2>&1; done
for c in 6 8 12 24 48; do RT_B200_SPARSE_CTAS=$c timeout 300 python tools/kstep_ab.py --envs 131072 --label "131072 envs sparse_ctas=$c" >> gpurun_out/u/sparse_ctas_sweep2.txt 2>&1; done
for c in 6 8 12 24; do RT_B200_SPARSE_CTAS=$c timeout 300 python tools/kstep_ab.py --envs 16384 --label "16384 envs sparse_ctas=$c" >> gpurun_out/u/sparse_ctas_sweep2.txt 2>&1; done
cat gpurun_out/u/sparse_ctas_sweep2.txt
