#!/bin/bash
# GPU session: rewritten smoother (bulk-copy stream + gather), fused V-cycle, one-sync GMRES
O=gpurun_out/r2_c
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -12 $O/pytest_gpu.log
for st in fp64 fp32 bf16; do
  OCMP_PATCH_STORAGE=$st timeout 300 python tools/kern_bench.py 128 2>&1 | tail -9 > $O/kern_$st.log; head -3 $O/kern_$st.log; tail -4 $O/kern_$st.log
done
timeout 300 python bench.py --no-cpu --steps 3 --warmup 2 > $O/bench2d_fp64.json 2> $O/bench2d_fp64.err
timeout 300 python bench.py --no-cpu --steps 3 --warmup 2 --precond-storage fp32 > $O/bench2d_fp32.json 2> $O/bench2d_fp32.err
timeout 300 python bench.py --workload ins3d_dim --N 48 --steps 2 --warmup 1 --no-cpu --precond-storage fp32 > $O/bench3d_fp32.json 2> $O/bench3d_fp32.err
for f in bench2d_fp64 bench2d_fp32 bench3d_fp32; do
python - <<PY
import json
try:
    l = json.loads(open('$O/$f.json').read().strip().splitlines()[-1])
    r = l['roofline']
    print('$f', 's/step %.4f' % l['value'], 'e2e %.4f' % l['e2e']['value'], 'its/step', l['problem']['gmres_its_per_step'], 'launches', l['gpu_launches'],
          'patch_apply %.0f GB/s (%.2f)' % (r['achieved'], r['frac']), 'spmv %.0f' % l['spmv_gbs'], l['kernel_time_share'])
except Exception as e:
    print('$f', 'FAILED', e); print(open('$O/$f.err').read()[-1500:])
PY
done
