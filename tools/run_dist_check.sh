#!/bin/bash
# 2-GPU check of the native distributed GMRES/V-cycle against the Python-driven one, then a timing at the bench size.
set -x
NP=${NP:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29511"
mkdir -p gpurun_out
OCMP_DIST_NATIVE=0 timeout 300 $TR tools/dist_ins.py 64 2>&1 | tail -3 | tee gpurun_out/dist_py_64.log
OCMP_DIST_NATIVE=1 timeout 300 $TR tools/dist_ins.py 64 2>&1 | tail -3 | tee gpurun_out/dist_native_64.log
OCMP_DIST_NATIVE=1 timeout 400 $TR tools/dist_ins.py ${NBIG:-256} 2>&1 | tail -3 | tee gpurun_out/dist_native_big.log
