#!/bin/bash
# everything profiles/ needs for round 2, in one gpurun call (1 GPU):  scripts/round2_final.sh [skip-tests]
mkdir -p gpurun_out
if [ -z "$1" ]; then ( time python -m pytest tests -m gpu -x -q ) 2>&1 | tail -5; fi
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -c 300 gpurun_out/r2_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2>&1
for k in de en longdoc; do python scripts/profile_one.py $((1<<30)) $k 2>&1 | tail -2; done | tee gpurun_out/r2_shapes.txt
python bench.py --steps 3 --warmup 3 --no-cpu --no-shapes --c5-slices 8 > gpurun_out/r2_bench_c5_share.json 2> gpurun_out/r2_bench_c5.err; tail -c 200 gpurun_out/r2_bench_c5.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-shapes > gpurun_out/ncu_l.log 2>&1; echo launches done
ncu --set full --import-source on --clock-control none -f -o gpurun_out/r2_full python scripts/profile_one.py $((1<<30)) > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log
ncu -i gpurun_out/r2_full.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_raw.csv 2>/dev/null
rm -f gpurun_out/r2_full.ncu-rep   # (hundreds of MB with sources; the raw page is what profiles/ keeps)
ls -la gpurun_out/r2_*
scripts/ncu_kernel.sh datok_b200/libdatok_b200.so walk "walk_fused_kernel.*1024"
scripts/ncu_kernel.sh datok_b200/libdatok_b200.so emit "compact_kernel<2>|compact_kernel<\(int\)2>"
rm -f gpurun_out/k_walk.ncu-rep gpurun_out/k_emit.ncu-rep
