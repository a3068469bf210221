#!/bin/bash
mkdir -p gpurun_out
for n in 32 48 64; do
  SECONDS=0
  timeout 900 python bench.py --workload ins3d_dim --N $n --steps 2 --warmup 1 --no-cpu > gpurun_out/bench3d_$n.json 2> gpurun_out/bench3d_$n.err
  echo "N=$n wall ${SECONDS}s rc=$?"; tail -c 1800 gpurun_out/bench3d_$n.json; tail -3 gpurun_out/bench3d_$n.err
done
