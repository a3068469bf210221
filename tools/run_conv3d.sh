#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python tools/conv_3d.py 32 1e-12
  OCMP_MG_NU=2 timeout 300 python tools/conv_3d.py 32 1e-12
  N0=4 timeout 300 python tools/conv_3d.py 32 1e-12
  LAM=4 timeout 300 python tools/conv_3d.py 32 1e-12 ) 2>&1 | grep -v Warning | tee gpurun_out/conv3d.log
