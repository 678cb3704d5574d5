#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pss_track -s 12 -c 1 -f -o $O/track_prof_c2 python bench.py --workload c2 > $O/r2_track_ncu_c2.log 2>&1
tail -2 $O/r2_track_ncu_c2.log | cut -c1-300
