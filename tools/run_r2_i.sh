#!/bin/bash
# 4 GPUs: native vs python cycle at small size, then the 1-GPU iteration counts at the totals the layout uses
for lay in sphere; do timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 tools/dist3d_check.py 24 $lay 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -3; done
CUDA_VISIBLE_DEVICES=0 timeout 200 python tools/conv_3d.py 40 1e-12 2>&1 | cut -c1-400 | tail -2 &
CUDA_VISIBLE_DEVICES=1 N0=2 timeout 500 python tools/conv_3d.py 80 1e-12 2>&1 | cut -c1-400 | tail -2 &
CUDA_VISIBLE_DEVICES=2 timeout 200 python tools/conv_3d.py 64 1e-12 2>&1 | cut -c1-400 | tail -2 &
wait
