#!/bin/bash
# Last GPU call of round 2: suite, smoke, default bench
O=gpurun_out/r2_last
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
timeout 600 python bench.py > $O/bench_1gpu.json 2> $O/bench_1gpu.err; tail -c 300 $O/bench_1gpu.err
python -c "
import json
l = json.loads(open('$O/bench_1gpu.json').read().strip().splitlines()[-1])
print('3d', l['value'], l['e2e']['value'], l['problem']['gmres_its_per_step'], sum(l['kernel_time_share'].values()))
t = l['ins2d']; print('2d', t['value'], t['e2e']['value'])
"
