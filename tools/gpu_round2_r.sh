#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pss_track -s 4 -c 1 -f -o $O/track_prof_r python bench.py --noise-only --no-e2e --no-spot-check --sustained-s 0 --no-alt --no-alone --steps 2 --pipeline serial > $O/r2_track_ncu_r.log 2>&1
tail -3 $O/r2_track_ncu_r.log
