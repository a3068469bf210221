#!/bin/bash
# GPU session: new bench.py (both arms), ncu captures of the inversion and contraction kernels
O=gpurun_out/r2_d
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
( time timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err ) 2> $O/bench_default.time
tail -c 400 $O/bench_default.err; cat $O/bench_default.time | tail -4
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err ) 2> $O/bench_ref.time
tail -c 400 $O/bench_ref.err; cat $O/bench_ref.time | tail -4
python - <<PY
import json
for f in ('bench_default', 'bench_ref'):
    try:
        l = json.loads(open('$O/%s.json' % f).read().strip().splitlines()[-1])
        keep = {k: v for k, v in l.items() if k not in ('config', 'ins2d')}
        print(f, json.dumps(keep)[:3000])
        if 'ins2d' in l:
            print('ins2d', json.dumps({k: v for k, v in l['ins2d'].items() if k != 'config'})[:2500])
    except Exception as e:
        print(f, 'FAILED', e)
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_patch_invert -c 2 -o $O/patch_invert -f python tools/kern_bench.py 64 > $O/ncu_invert.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_contract -c 3 -o $O/contract -f python tools/kern_bench.py 64 > $O/ncu_contract.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_patch_apply_stream -c 2 -o $O/patch_apply_stream -f python tools/kern_bench.py 128 > $O/ncu_apply.log 2>&1
ls -la $O
