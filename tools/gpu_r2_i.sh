#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2i_launches_pair_b1.csv python tools/profile_step.py 1 1 1 > gpurun_out/r2i_ncu.log 2>&1; echo "ncu pair b1 exit $?"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2i_launches_pair_b8.csv python tools/profile_step.py 8 1 0 > gpurun_out/r2i_ncu8.log 2>&1; echo "ncu pair b8 exit $?"
EGR_TC_NO_PAIR=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2i_launches_nopair_b8.csv python tools/profile_step.py 8 1 0 > gpurun_out/r2i_ncu8n.log 2>&1; echo "ncu nopair b8 exit $?"
timeout 300 python tools/section_times.py 1 1 2>/dev/null | tail -7
EGR_TC_NO_PAIR=1 timeout 300 python tools/section_times.py 1 1 2>/dev/null | tail -7
