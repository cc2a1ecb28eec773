mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "general_component or dropin or f32io or kf_gain or pogox or pipeline_equals" > gpurun_out/gputests_r2b.txt 2>&1; tail -4 gpurun_out/gputests_r2b.txt
./tools/_build/tune_solve 65536 20 30 > gpurun_out/tune_pw.txt 2>&1
./tools/_build/tune_solve 524288 20 6 > gpurun_out/tune_pw_big.txt 2>&1
python bench.py --steps 100 --warmup 5 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; tail -c 600 gpurun_out/bench_r2a.err
python tools/tick_probe.py 0 > gpurun_out/tick_default.txt 2>&1
DEKF_B200_SO=$PWD/tools/_build/libdekf_legrolled.so python tools/tick_probe.py 0 > gpurun_out/tick_legrolled.txt 2>&1
cat gpurun_out/tune_pw.txt gpurun_out/tune_pw_big.txt gpurun_out/tick_default.txt gpurun_out/tick_legrolled.txt
