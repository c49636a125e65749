#!/bin/bash
# source-level counters of one launch of a kernel: scripts/ncu_kernel.sh lib.so tag kernel-regex [bytes] [kind] [launch-skip]
lib=$1; tag=$2; kre=$3; size=${4:-268435456}; kind=${5:-de}; skip=${6:-1}
DATOK_B200_LIB=$PWD/$lib ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --import-source on --clock-control none \
  --kernel-name-base demangled -k "regex:$kre" --launch-skip $skip --launch-count 1 -f -o gpurun_out/k_$tag python scripts/profile_one.py $size $kind > gpurun_out/ncu_k_$tag.log 2>&1
tail -2 gpurun_out/ncu_k_$tag.log
ncu -i gpurun_out/k_$tag.ncu-rep --page source --csv > gpurun_out/k_${tag}_src.csv 2>/dev/null
ncu -i gpurun_out/k_$tag.ncu-rep --page raw --csv > gpurun_out/k_${tag}_raw.csv 2>/dev/null
