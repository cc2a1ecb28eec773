#!/bin/bash
# ncu captures of the per-tick kernels of the incremental path (one launch each, steady state)
set -e
mkdir -p gpurun_out
for k in k_ekf k_assemble k_solve_incr; do
  ncu --set full --clock-control none --import-source on -k regex:"^${k}\$|${k}<" -s 40 -c 1 -f -o gpurun_out/prof_${k}_r01 \
      python tools/tick_probe.py 1 > gpurun_out/ncu_${k}.log 2>&1 || true
  ncu -i gpurun_out/prof_${k}_r01.ncu-rep --page raw --csv > gpurun_out/prof_${k}_r01_raw.csv 2>/dev/null || true
done
ls -la gpurun_out | tail
