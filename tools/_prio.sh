(
python tools/run_probe.py 0 200
python tools/run_probe.py 1 200
python tools/tick_probe.py 0
) 2>&1 | grep "run_probe\|window_solve" | tee gpurun_out/asm_hoist_probe.txt
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
