#!/bin/bash
# usage: tools/r02b_multigpu.sh N   (inside gpurun --gpus N): end-of-round-2 scaling line and config-5 line (profiles/r02b_*)
N=$1
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
$RUN bench.py --gpus $N --steps 100 --warmup 5 --no-configs > gpurun_out/r02b_bench_fp64_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
tail -c 300 gpurun_out/bench_${N}gpu.err
PER=$((1000000 / N))
$RUN bench.py --gpus $N --steps 24 --warmup 3 --instances $PER --N 100 --no-configs --no-cpu-baseline --e2e-steps 8 > gpurun_out/r02b_bench_cfg5_1M_N100_${N}gpu.json 2> gpurun_out/bench_cfg5_${N}gpu.err
tail -c 300 gpurun_out/bench_cfg5_${N}gpu.err
python - <<PY
import json
for f in ("gpurun_out/r02b_bench_fp64_${N}gpu.json", "gpurun_out/r02b_bench_cfg5_1M_N100_${N}gpu.json"):
    try:
        d = json.loads([x for x in open(f) if x.startswith("{")][-1])
        print(f, "value %.3e ms %.4f incr %.3e e2e %.3e (h2d %.1f GB/s/rank) e2e64 %.3e step_frac %.3f" % (d["value"], d["ms_per_step"], d["incremental"]["value"], d["e2e"]["value"], d["e2e"]["h2d_gbs"] / d["n_gpus"], d["e2e_f64_io"]["value"], d["roofline"]["step"]["frac"]))
    except Exception as e:
        print(f, "failed", e)
PY
