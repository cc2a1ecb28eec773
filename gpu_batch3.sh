mkdir -p gpurun_out
python tools/_dbg_box.py > gpurun_out/dbg_box.txt 2>&1; tail -12 gpurun_out/dbg_box.txt
python -m pytest tests -m gpu -q -x -k "general_component or pogox" > gpurun_out/gputests_r2d.txt 2>&1; tail -3 gpurun_out/gputests_r2d.txt
for ch in 8 8 4 16; do
  DEKF_HOST_CHUNK=$ch python bench.py --steps 60 --warmup 5 --no-configs --no-cpu-baseline --e2e-steps 60 > gpurun_out/bench_e2e_ch$ch.json 2>/dev/null
  python - <<PY
import json
d=json.loads([x for x in open('gpurun_out/bench_e2e_ch$ch.json') if x.startswith('{')][-1])
print('chunk $ch value %.3e e2e %.3e (h2d %.1f d2h %.1f GB/s) e2e64 %.3e' % (d['value'], d['e2e']['value'], d['e2e']['h2d_gbs'], d['e2e']['d2h_gbs'], d['e2e_f64_io']['value']))
PY
done
