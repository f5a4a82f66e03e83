#!/bin/bash
# Lean ncu capture through the plain-C driver (no Python under the profiler).  Usage: tools/profile_c.sh <tag> [driver args]
set -e
tag=$1; shift
mkdir -p gpurun_out
drv=tools/c_driver/amh_c_driver
$drv "$@" | tee gpurun_out/drv_${tag}.txt
ncu --set full --clock-control none --import-source on -k regex:'step|sweep|ram_warp' -s 3 -c 1 -f -o gpurun_out/prof_${tag} $drv "$@" > gpurun_out/ncu_full_${tag}.log 2>&1 || tail -5 gpurun_out/ncu_full_${tag}.log
ncu -i gpurun_out/prof_${tag}.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2>/dev/null || true
ncu -i gpurun_out/prof_${tag}.ncu-rep --page source --csv --print-source sass > gpurun_out/prof_${tag}_src.csv 2>/dev/null || true
ls -la gpurun_out | grep ${tag}
