#!/bin/bash
# everything the round's profiles/ need, in one gpurun call:  scripts/round_final.sh tag [old_lib.so]
tag=$1; old=$2
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
[ -n "$old" ] && scripts/ab_variants.sh $((1<<30)) de $old
for k in de en longdoc; do python scripts/profile_one.py $((1<<30)) $k 2>&1 | tail -1; done | tee gpurun_out/shapes_$tag.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$tag.log 2> gpurun_out/bench_$tag.err; tail -c 600 gpurun_out/bench_$tag.log
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_$tag.log 2>&1; tail -c 300 gpurun_out/bench_ref_$tag.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_l_$tag.log 2>&1; echo launches done
ncu --set full --import-source on --clock-control none -f -o gpurun_out/prof_$tag python scripts/profile_one.py $((1<<30)) > gpurun_out/ncu_$tag.log 2>&1; tail -1 gpurun_out/ncu_$tag.log
