timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tools/check_p2p.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -6
LFS_P2P_MULTICAST=0 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 tools/check_p2p.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -4
