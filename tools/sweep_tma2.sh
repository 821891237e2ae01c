#!/bin/bash
run() { python tools/kbench.py --reps 20 --moments-mb 64 2>&1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('raster %.4f ms %.3f | k_step %.4f | group %.4f ms -> %.1f M agent-steps/s' % (d['k_raster_vec']['ms'], d['k_raster_vec']['frac'], d['k_step']['ms'], d['step_group']['ms'], d['step_group']['agent_steps_per_s']/1e6))"; }
for ctas in 1 2 3 4; do for st in 4 8; do echo "TMA raster standalone (serial step) base_ctas=$ctas stages=$st"; RT_B200_OVERLAP=0 RT_B200_RASTER_TMA=1 RT_B200_BASE_CTAS=$ctas RT_B200_TMA_STAGES=$st run; done; done
