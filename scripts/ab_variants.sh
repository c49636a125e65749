#!/bin/bash
# A/B of library builds on one box: scripts/ab_variants.sh bytes kind lib1.so lib2.so ...
# (each through scripts/profile_one.py; prints the second, warm, transduction)
size=$1; kind=$2; shift; shift
for lib in "$@"; do
  echo "== $lib $kind $size"
  DATOK_B200_LIB=$PWD/$lib python scripts/profile_one.py $size $kind 2>&1 | tail -2 | head -1
done
