timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_ssim|k_blend|k_adam|k_rs_scatter|k_emit|k_preprocess" -s 30 -c 14 -o gpurun_out/prof_r01g -f python tools/profile_view.py C3 2 > gpurun_out/prof_r01g.log 2>&1
tail -1 gpurun_out/prof_r01g.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r01d.csv python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/b.log 2>&1
tail -1 gpurun_out/b.log | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -1 gpurun_out/bench_final.json | cut -c1-400
