#!/bin/bash
# A/B sweep of the overlapped-step knobs on one GPU (prints compact lines)
run() { python tools/kbench.py --reps 20 --moments-mb 64 2>&1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('raster %.4f ms %.3f | k_step %.4f | group %.4f ms -> %.1f M agent-steps/s' % (d['k_raster_vec']['ms'], d['k_raster_vec']['frac'], d['k_step']['ms'], d['step_group']['ms'], d['step_group']['agent_steps_per_s']/1e6))"; }
echo "serial"; RT_B200_OVERLAP=0 run
for ctas in 1 2 3 4 8; do for u in 2 4; do echo "overlap base_ctas=$ctas unroll=$u"; RT_B200_BASE_CTAS=$ctas RT_B200_BASE_UNROLL=$u run; done; done
for epb in 16 32; do echo "overlap epb=$epb base 2x4"; RT_B200_EPB=$epb run; done
for ctas in 8 16 32; do for u in 1 2 4; do echo "full raster ctas=$ctas unroll=$u"; RT_B200_OVERLAP=0 RT_B200_RASTER_CTAS=$ctas RT_B200_RASTER_UNROLL=$u run; done; done
