#!/bin/bash
# Default bench after the coarse-level penalty fix (multigrid.inherit_cell_penalty), then the GPU suite as far as the
# remaining box time allows
O=gpurun_out/r2_penalty
mkdir -p $O
timeout 130 python bench.py > $O/bench_1gpu.json 2> $O/bench_1gpu.err; echo "bench rc=$?"; tail -c 300 $O/bench_1gpu.err
python -c "
import json
l = json.loads(open('$O/bench_1gpu.json').read().strip().splitlines()[-1])
print('3d', l['value'], l['e2e']['value'], l['problem']['gmres_its_per_step'], l['problem']['gmres_its_per_solve'][:2], sum(l['kernel_time_share'].values()))
t = l['ins2d']; print('2d', t['value'], t['e2e']['value'], t['problem']['gmres_its_per_step'])
"
timeout 45 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
