#!/bin/bash
run() { python tools/kbench.py --reps 20 --moments-mb 64 2>&1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('raster %.4f ms | k_step %.4f | group %.4f ms -> %.1f M agent-steps/s' % (d['k_raster_vec']['ms'], d['k_step']['ms'], d['step_group']['ms'], d['step_group']['agent_steps_per_s']/1e6))"; }
for h in 19 3; do for c in 1 2 4; do echo "sparse_lsu dynamic hints=$h ctas=$c"; RT_B200_SPARSE_HINTS=$h RT_B200_SPARSE_CTAS=$c run; done; done
