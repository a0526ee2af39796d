timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 300 python bench.py --steps 8 --warmup 3 --no-extras > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_x.json").read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["e2e"]["value"],1), d["config"]["instances_per_view"], {k:round(x,3) for k,x in d["stage_ms_per_view"].items()})
PY
