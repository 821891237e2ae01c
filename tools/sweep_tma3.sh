#!/bin/bash
run() { python tools/kbench.py --reps 20 --moments-mb 64 2>&1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('raster %.4f ms %.3f | k_step %.4f | group %.4f ms -> %.1f M agent-steps/s' % (d['k_raster_vec']['ms'], d['k_raster_vec']['frac'], d['k_step']['ms'], d['step_group']['ms'], d['step_group']['agent_steps_per_s']/1e6))"; }
for co in 100 75; do
echo "== carveout $co"
echo "serial"; RT_B200_CARVEOUT=$co RT_B200_OVERLAP=0 run
for ctas in 1 2; do for st in 4 8; do echo "overlap TMA base_ctas=$ctas stages=$st"; RT_B200_CARVEOUT=$co RT_B200_OVERLAP=1 RT_B200_BASE_CTAS=$ctas RT_B200_TMA_STAGES=$st run; done; done
for ctas in 2 4 6; do echo "overlap LSU base_ctas=$ctas unroll=4"; RT_B200_CARVEOUT=$co RT_B200_OVERLAP=1 RT_B200_BASE_TMA=0 RT_B200_BASE_CTAS=$ctas RT_B200_BASE_UNROLL=4 run; done
done
