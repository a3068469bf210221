#!/bin/bash
# Record of a measurement: OCMP_PATCH_CHUNK / OCMP_PATCH_CVT were temporary switches of ocmp_patch.cu (chunk size of the
# bulk copies; FP32 -> FP64 widening on the XU / integer pipe). The measured best (8 KB, F2F) is compiled in, the switches
# are gone; results: profiles/r2_patch_chunk_sweep.log, profiles/r2_patch_cvt_sweep.log.
# smoother products: chunk size of the bulk copies (bytes per stage) against the per-chunk synchronisation cost
O=gpurun_out/r2_cvt
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
for st in fp32 bf16 fp64; do for m in 4096 6144; do
  echo "storage $st chunk $m" | tee -a $O/chunk.log
  OCMP_PATCH_CVT=0 OCMP_PATCH_CHUNK=$m OCMP_PATCH_STORAGE=$st timeout 300 python tools/kern_bench.py 128 2>&1 | grep "level [45]" | tee -a $O/chunk.log
done; done
