#!/bin/bash
# ncu captures of the dominant kernel (run under gpurun, one GPU).  Usage: tools/profile_k1.sh <tag> [bench args...]
# Writes gpurun_out/prof_<tag>.ncu-rep (+ raw/source CSV) and gpurun_out/launches_<tag>.csv
set -e
tag=$1; shift
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 "$@" > gpurun_out/ncu_launch_${tag}.log 2>&1 || true
ncu --set full --clock-control none --import-source on -k regex:'mh_step|mala_step|ram_step|stretch_sweep' -s 3 -c 1 -f -o gpurun_out/prof_${tag} \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 0 --mcmc-steps-per-launch 20 "$@" > gpurun_out/ncu_full_${tag}.log 2>&1 || true
ncu -i gpurun_out/prof_${tag}.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2>/dev/null || true
ncu -i gpurun_out/prof_${tag}.ncu-rep --page source --csv --print-source sass > gpurun_out/prof_${tag}_src.csv 2>/dev/null || true
ls -la gpurun_out | grep ${tag}
