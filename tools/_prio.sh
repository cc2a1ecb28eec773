(
python tools/run_probe.py 0 200
DEKF_NO_ASM_SPLIT=1 python tools/run_probe.py 0 200
python tools/run_probe.py 1 200
) 2>&1 | grep run_probe | tee gpurun_out/asm_split_probe.txt
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pipeline or run_host" 2>&1 | tail -5
