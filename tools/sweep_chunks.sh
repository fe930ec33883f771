#!/bin/bash
# Frame time against the sample-chunk schedule of k_render's work list (B200RT_CHUNKS levels, sizes falling by
# B200RT_CHUNK_RATIO per level) at a given spp; the per-GPU share of a multi-GPU sample split is a low-spp frame.
# Usage: tools/sweep_chunks.sh <scene> <spp> "<chunk counts>" "<ratios>"
for c in $3; do for q in ${4:-1}; do
  echo -n "spp=$2 CHUNKS=$c RATIO=$q  "
  B200RT_CHUNKS=$c B200RT_CHUNK_RATIO=$q timeout 120 python tools/profile_frame.py $2 $1 | tail -1
done; done
