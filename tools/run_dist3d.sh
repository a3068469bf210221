#!/bin/bash
NP=${NP:-2}; N=${N:-32}
mkdir -p gpurun_out
SECONDS=0
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $NP --workload ins3d_dim --N $N --steps 2 --warmup 1 > gpurun_out/bench3d_${N}_${NP}gpu.json 2> gpurun_out/bench3d_${N}_${NP}gpu.err
echo "rc=$? wall ${SECONDS}s"; tail -c 2500 gpurun_out/bench3d_${N}_${NP}gpu.json; grep -v Warning gpurun_out/bench3d_${N}_${NP}gpu.err | tail -5
