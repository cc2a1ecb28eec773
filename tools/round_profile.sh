#!/bin/bash
# One GPU-box call that produces everything profiles/ holds for the round (bench lines, launch list, ncu captures).
mkdir -p gpurun_out
python bench.py --steps 200 --warmup 10 > gpurun_out/r01_bench_fp64_1gpu_v3.json 2> gpurun_out/bench_v3.err
python bench.py --steps 100 --warmup 5 --precision fp32 --no-cpu-baseline > gpurun_out/r01_bench_fp32_1gpu_v3.json 2>> gpurun_out/bench_v3.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01_bench_reference_arm.json 2>> gpurun_out/bench_v3.err
python tools/variants_bench.py > gpurun_out/r01_variants.jsonl 2> gpurun_out/variants.err
# launch list of the bench command (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_ekf|k_assemble|k_solve|k_kf|k_fused|k_init|k_fma|k_get|k_vo|k_resweep|k_arrival" \
    -c 700 --csv --log-file gpurun_out/r01_launches_v3.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 4 > gpurun_out/ncu_launches.log 2>&1
for k in k_solve_tma k_solve_incr_tma k_solve_incr k_assemble k_ekf; do
  mode=1; [ "$k" = "k_solve_tma" ] && mode=0
  skip=45; [ "$k" = "k_solve_incr_tma" ] && skip=5
  ncu --set full --clock-control none --import-source on -k regex:"^${k}\$|${k}<" -s $skip -c 1 -f -o gpurun_out/prof_${k}_r01v3 \
      python tools/tick_probe.py $mode > gpurun_out/ncu_${k}.log 2>&1
  ncu -i gpurun_out/prof_${k}_r01v3.ncu-rep --page raw --csv > gpurun_out/prof_${k}_r01v3_raw.csv 2>/dev/null
done
python tools/tick_probe.py 1 > gpurun_out/tick_probe_incr.txt 2>&1
python tools/tick_probe.py 0 > gpurun_out/tick_probe_full.txt 2>&1
ls -la gpurun_out | tail -30
