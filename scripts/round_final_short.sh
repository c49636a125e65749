#!/bin/bash
# tests + bench lines + launch list (no full ncu capture):  scripts/round_final_short.sh tag
tag=$1
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for k in de en longdoc; do python scripts/profile_one.py $((1<<30)) $k 2>&1 | tail -1; done | tee gpurun_out/shapes_$tag.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$tag.log 2> gpurun_out/bench_$tag.err; tail -c 400 gpurun_out/bench_$tag.log
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_$tag.log 2>&1; tail -c 200 gpurun_out/bench_ref_$tag.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_l_$tag.log 2>&1; echo launches done
