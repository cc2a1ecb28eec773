mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/gputests_r2c.txt 2>&1; tail -4 gpurun_out/gputests_r2c.txt
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_r2.json 2> gpurun_out/bench_ref_r2.err
python bench.py --steps 200 --warmup 10 > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; tail -c 400 gpurun_out/bench_r2b.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2.txt 2>&1; tail -5 gpurun_out/smoke_r2.txt
