timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 tools/check_p2p.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -4
LFS_P2P_MULTICAST=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 tools/check_p2p.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 8 --warmup 3 --no-extras > gpurun_out/bench_4gpu_auto.json 2> gpurun_out/bench_4gpu_auto.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_4gpu_auto.json").read().strip().splitlines()[-1])
    print("auto", round(d["value"],1), round(d["ms_per_step"],3), round(d["e2e"]["value"],1), d["config"]["parallelism"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_4gpu_auto.err").read()[-1500:])
PY
