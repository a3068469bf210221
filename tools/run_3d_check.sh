#!/bin/bash
# 1-GPU: new 3-D parity tests, then the 3-D INS-DIM bench at two sizes
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "3d or dim" 2>&1 | tail -15 | tee gpurun_out/test3d.log
timeout 600 python bench.py --workload ins3d_dim --N 16 --steps 2 --warmup 1 --no-cpu 2>&1 | tail -3 | tee gpurun_out/bench3d_16.json
timeout 900 python bench.py --workload ins3d_dim --N 32 --steps 2 --warmup 1 --no-cpu 2>&1 | tail -3 | tee gpurun_out/bench3d_32.json
nvidia-smi --query-gpu=memory.used --format=csv
