set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for o in "pre_bwd_split=0" "pre_bwd_split=1"; do
LFS_OPTIONS="$o" timeout 300 python bench.py --steps 8 --warmup 3 --no-extras > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_x.json").read().strip().splitlines()[-1])
print("$o", round(d["value"],1), {k:round(x,3) for k,x in d["stage_ms_per_view"].items()})
PY
done
