#!/bin/bash
for m in 8 12 16 24 32; do
  echo "== EGR_TC_SPLIT_MIN=$m"
  EGR_TC_SPLIT_MIN=$m timeout 300 python tools/section_times.py 1 1 2>/dev/null | grep "unet\|whole"
  EGR_TC_SPLIT_MIN=$m timeout 300 python tools/section_times.py 8 1 2>/dev/null | grep "unet\|whole"
done
