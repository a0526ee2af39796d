for v in 0 1 2; do
LFS_OPTIONS="fwd_variant=$v" timeout 300 python bench.py --steps 8 --warmup 3 --no-extras > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_x.json").read().strip().splitlines()[-1])
print("fwd_variant $v", round(d["value"],1), {k:round(x,3) for k,x in d["stage_ms_per_view"].items()})
PY
done
