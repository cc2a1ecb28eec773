(
python tools/run_probe.py 0 200
for w in 2 3 4 5 6 8; do DEKF_SPLIT_WAYS=$w python tools/run_probe.py 0 200; done
for w in 3 4 8; do DEKF_PRIO=2 DEKF_SPLIT_WAYS=$w python tools/run_probe.py 0 200; done
DEKF_SPLIT_TILES=148 python tools/run_probe.py 0 200
DEKF_SPLIT_TILES=256 python tools/run_probe.py 0 200
DEKF_SPLIT_TILES=364 python tools/run_probe.py 0 200
) 2>&1 | grep run_probe | tee gpurun_out/ways_probe.txt
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pipeline or run_host" 2>&1 | tail -5
