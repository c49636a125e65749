#!/bin/bash
# source-level counters of the first fused walk launch: scripts/ncu_walk.sh lib.so tag [bytes] [kind]
lib=$1; tag=$2; size=${3:-268435456}; kind=${4:-de}
DATOK_B200_LIB=$PWD/$lib ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --import-source on --clock-control none \
  --kernel-name-base demangled -k "regex:walk_fused_kernel.*1024" --launch-skip 1 --launch-count 1 -f -o gpurun_out/walk_$tag python scripts/profile_one.py $size $kind > gpurun_out/ncu_walk_$tag.log 2>&1
tail -2 gpurun_out/ncu_walk_$tag.log
