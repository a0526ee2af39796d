timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
timeout 300 python tools/bench_fastgs.py C3 4 3 2>&1 | tail -1 | cut -c1-400
timeout 300 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
tail -1 gpurun_out/bench_full.json | cut -c1-300
