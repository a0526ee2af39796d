set -x
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 --no-extras > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
tail -1 gpurun_out/bench_2gpu.json | cut -c1-900
tail -5 gpurun_out/bench_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_2gpu_ref.json 2> gpurun_out/bench_2gpu_ref.err
tail -1 gpurun_out/bench_2gpu_ref.json | cut -c1-600
