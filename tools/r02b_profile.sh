#!/bin/bash
# One GPU-box call: end-of-round-2 bench lines and ncu captures (profiles/r02b_*).  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
python bench.py --steps 200 --warmup 10 > gpurun_out/r02b_bench_fp64_1gpu.json 2> gpurun_out/r02b_bench.err
tail -c 300 gpurun_out/r02b_bench.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02b_bench_reference_arm.json 2>> gpurun_out/r02b_bench.err
# launch list of the bench command (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_ekf|k_assemble|k_solve|k_kf|k_fused|k_init|k_fma|k_get|k_vo|k_zero|k_resweep|k_arrival|k_widen|k_narrow|k_box|k_foot" \
    -c 900 --csv --log-file gpurun_out/r02b_launches.csv \
    python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-configs --e2e-steps 4 > gpurun_out/ncu_launches.log 2>&1
for k in k_solve_tma k_assemble k_ekf; do
  ncu --set full --clock-control none --import-source on -k regex:"^${k}\$|${k}<" -s 45 -c 1 -f -o gpurun_out/r02b_prof_${k} \
      python tools/tick_probe.py 0 > gpurun_out/ncu_${k}.log 2>&1
  ncu -i gpurun_out/r02b_prof_${k}.ncu-rep --page raw --csv > gpurun_out/r02b_${k}_ncu_raw.csv 2>/dev/null
  rm -f gpurun_out/r02b_prof_${k}.ncu-rep
done
ncu --set full --clock-control none --import-source on -k regex:"k_fused_roles" -s 80 -c 1 -f -o gpurun_out/r02b_prof_k_fused_roles \
    python tools/b1_probe.py 0 > gpurun_out/ncu_k_fused_roles.log 2>&1
ncu -i gpurun_out/r02b_prof_k_fused_roles.ncu-rep --page raw --csv > gpurun_out/r02b_k_fused_roles_b1_ncu_raw.csv 2>/dev/null
rm -f gpurun_out/r02b_prof_k_fused_roles.ncu-rep
(python tools/b1_probe.py 0 300; python tools/b1_probe.py 1 300; python tools/tick_probe.py 0; python tools/tick_probe.py 1; python tools/tick_probe.py 0 fp64 ragged
 python tools/run_probe.py 0 200; python tools/run_probe.py 1 200; python tools/run_probe.py 0 120 65536 ragged; DEKF_VO_COMPACT=1 python tools/run_probe.py 0 120 65536 ragged
 DEKF_PRIO=1 DEKF_NO_ASM_SPLIT=1 python tools/run_probe.py 0 200) 2>&1 | grep "batch-\|window_solve\|run_probe" | tee gpurun_out/r02b_probes.txt
python - <<'PY'
import json
d = json.loads([x for x in open("gpurun_out/r02b_bench_fp64_1gpu.json") if x.startswith("{")][-1])
r = d["roofline"]
print("value %.4e ms %.4f e2e %.3e e2e64 %.3e b1 %.1f/%.1f us host %.1f/%.1f roofline %s %.3f step %.3f ragged %.3e" % (
    d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e_f64_io"]["value"], d["latency_batch1"]["device_us_per_tick"],
    d["incremental"]["latency_batch1"]["device_us_per_tick"], d["latency_batch1"]["host_call_us_median"],
    d["incremental"]["latency_batch1"]["host_call_us_median"], r["bound"], r["frac"], r["step"]["frac"], d["vo_ragged"]["full"]["value"]))
print({k: (v.get("value"), v.get("ms_per_step")) for k, v in d.get("configs", {}).items()})
ref = json.loads([x for x in open("gpurun_out/r02b_bench_reference_arm.json") if x.startswith("{")][-1])
print("reference arm", ref["value"], ref["cpu_baseline"]["sample"][:120], "cpu_baseline", d["cpu_baseline"]["value"])
PY
