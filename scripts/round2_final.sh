#!/bin/bash
# everything profiles/ needs for round 2, in one gpurun call (1 GPU):  scripts/round2_final.sh
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -c 300 gpurun_out/r2_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2>&1
for k in de en longdoc; do python scripts/profile_one.py $((1<<30)) $k 2>&1 | tail -1; done | tee gpurun_out/r2_shapes.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-shapes > gpurun_out/ncu_l.log 2>&1; echo launches done
ncu --set full --import-source on --clock-control none -f -o gpurun_out/r2_full python scripts/profile_one.py $((1<<30)) > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log
ls -la gpurun_out/r2_*
