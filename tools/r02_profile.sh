#!/bin/bash
# One GPU-box call: round-2 ncu captures (profiles/r02_*).  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
# launch list of the bench command (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_ekf|k_assemble|k_solve|k_kf|k_fused|k_init|k_fma|k_get|k_vo|k_resweep|k_arrival|k_widen|k_narrow|k_box|k_foot" \
    -c 900 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-configs --e2e-steps 4 > gpurun_out/ncu_launches.log 2>&1
for k in k_solve_tma k_solve_incr_tma k_solve_incr k_assemble k_ekf; do
  mode=1; [ "$k" = "k_solve_tma" ] && mode=0
  skip=45; [ "$k" = "k_solve_incr_tma" ] && skip=5
  ncu --set full --clock-control none --import-source on -k regex:"^${k}\$|${k}<" -s $skip -c 1 -f -o gpurun_out/r02_prof_${k} \
      python tools/tick_probe.py $mode > gpurun_out/ncu_${k}.log 2>&1
  ncu -i gpurun_out/r02_prof_${k}.ncu-rep --page raw --csv > gpurun_out/r02_${k}_ncu_raw.csv 2>/dev/null
  ncu -i gpurun_out/r02_prof_${k}.ncu-rep --page source --csv > gpurun_out/r02_${k}_ncu_source.csv 2>/dev/null
  rm -f gpurun_out/r02_prof_${k}.ncu-rep   # gpurun copies back at most 64 MiB
done
ncu --set full --clock-control none --import-source on -k regex:"k_fused" -s 80 -c 1 -f -o gpurun_out/r02_prof_k_fused_b1 \
    python tools/b1_probe.py 0 > gpurun_out/ncu_k_fused.log 2>&1
ncu -i gpurun_out/r02_prof_k_fused_b1.ncu-rep --page raw --csv > gpurun_out/r02_k_fused_b1_ncu_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_prof_k_fused_b1.ncu-rep --page source --csv > gpurun_out/r02_k_fused_b1_ncu_source.csv 2>/dev/null
rm -f gpurun_out/r02_prof_k_fused_b1.ncu-rep
python tools/b1_probe.py 0 > gpurun_out/b1_full.txt 2>&1
python tools/b1_probe.py 1 > gpurun_out/b1_incr.txt 2>&1
python tools/tick_probe.py 0 > gpurun_out/tick_probe_full.txt 2>&1
python tools/tick_probe.py 1 > gpurun_out/tick_probe_incr.txt 2>&1
cat gpurun_out/b1_full.txt gpurun_out/b1_incr.txt gpurun_out/tick_probe_full.txt gpurun_out/tick_probe_incr.txt
ls -la gpurun_out | tail -20
