#!/bin/bash
run() { python tools/kbench.py --reps 20 --moments-mb 64 2>&1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('raster %.4f ms | k_step %.4f | group %.4f ms -> %.1f M agent-steps/s' % (d['k_raster_vec']['ms'], d['k_step']['ms'], d['step_group']['ms'], d['step_group']['agent_steps_per_s']/1e6))"; }
for m in 4 3; do
  RT_B200_NVCC_EXTRA="-DRT_STEP_MIN_CTAS=$m" python -m rad_team_b200.build --force > /dev/null
  grep -E "k_stepE" -A2 rad_team_b200/build.log | grep -E "Used"
  echo "min_ctas=$m"; run
done
python -m rad_team_b200.build --force > /dev/null
