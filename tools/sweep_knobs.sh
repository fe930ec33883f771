#!/bin/bash
# Sweep the scheduler thresholds of k_render (B200RT_BATCH x B200RT_FRAC8) on one scene.
# Usage: tools/sweep_knobs.sh <scene> <spp> "<batch values>" "<frac8 values>"
scene=$1; spp=$2
for b in $3; do for f in $4; do
  echo -n "BATCH=$b FRAC8=$f  "
  B200RT_BATCH=$b B200RT_FRAC8=$f timeout 120 python tools/profile_frame.py $spp $scene | tail -1
done; done
