#!/bin/bash
O=gpurun_out/r2_g
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -12 $O/pytest_gpu.log
OCMP_INVERT_512=0 OCMP_PATCH_STORAGE=fp32 timeout 300 python tools/kern_bench.py 128 2>&1 | grep 'level 5\|pre.Update'
OCMP_PATCH_STORAGE=fp32 timeout 300 python tools/kern_bench.py 128 2>&1 | tail -9
timeout 600 python bench.py --no-cpu > $O/bench_default.json 2> $O/bench_default.err; tail -c 300 $O/bench_default.err
python - <<PY
import json
l = json.loads(open('$O/bench_default.json').read().strip().splitlines()[-1])
print('its', l['problem']['gmres_its_per_solve'], l['problem']['gmres_its_per_solve_e2e']); print('3d', l['value'], l['e2e']['value'], l['problem']['gmres_its_per_step'], l['gpu_launches'], l['kernel_time_share'])
print('   asm', l['roofline_assembly'])
t = l['ins2d']
print('2d', t['value'], t['e2e']['value'], t['problem']['gmres_its_per_step'], t['gpu_launches'], t['kernel_time_share'])
print('   asm', t['roofline_assembly'])
PY

