#!/bin/bash
run() { python tools/kbench.py --reps 20 --moments-mb 64 2>&1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('raster %.4f ms | k_step %.4f | group %.4f ms -> %.1f M agent-steps/s' % (d['k_raster_vec']['ms'], d['k_step']['ms'], d['step_group']['ms'], d['step_group']['agent_steps_per_s']/1e6))"; }
for epb in 64 32 16; do echo "epb=$epb"; RT_B200_EPB=$epb run; done
