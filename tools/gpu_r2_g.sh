#!/bin/bash
mkdir -p gpurun_out
for v in "" "EGR_MEGA_NO_PREFETCH=1" "EGR_MEGA_GN_TWO_OPS=1" "EGR_MEGA_NO_DEFER=1"; do
  echo "=== variant: $v"
  env $v timeout 300 python tools/mega_trace.py 1 1 2>&1 | grep -v Warning | head -n 22 | tail -n 20
done
