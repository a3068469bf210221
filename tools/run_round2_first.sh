#!/bin/bash
# First GPU session of round 2 (one GPU): everything written after round 1's GPU budget ran out, in one call.
#   gpurun --timeout 1500 -- 'bash tools/run_round2_first.sh'
# Outputs under gpurun_out/r2_first/ (scratch; copy what is to be judged into profiles/).
O=gpurun_out/r2_first
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
# 1. all GPU tests, no -x: the late additions (tests/test_zz_gpu_late_additions.py) have never run on a B200
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
tail -15 $O/pytest_gpu.log
# 2. bench: round-1 configuration, coarse levels reused (new default), FP32-stored preconditioner data
timeout 300 python bench.py --full-mg-setup --no-cpu > $O/bench_full_setup.json 2> $O/bench_full_setup.err
timeout 300 python bench.py --no-cpu > $O/bench_default.json 2> $O/bench_default.err
timeout 300 python bench.py --no-cpu --precond-storage fp32 > $O/bench_fp32.json 2> $O/bench_fp32.err
timeout 300 python bench.py --no-cpu --precond-storage bf16 > $O/bench_bf16.json 2> $O/bench_bf16.err
timeout 300 python bench.py --no-cpu --precond-storage bf16 --lag-smoother > $O/bench_all.json 2> $O/bench_all.err
for f in full_setup default fp32 bf16 all; do
  python - <<PY
import json
try:
    l = json.loads(open('$O/bench_$f.json').read().strip().splitlines()[-1])
    r = l['roofline']
    print('$f', 's/step %.4f' % l['value'], 'e2e %.4f' % l['e2e']['value'], 'its/step', l['problem']['gmres_its_per_step'],
          'patch_apply %.0f GB/s (%.2f)' % (r['achieved'], r['frac']), 'spmv %.0f' % l['spmv_gbs'], l['kernel_time_share'],
          'matrix-free apply', l.get('matrix_free_apply'), 'csr spmv ms', l['roofline_spmv']['avg_launch_ms'])
except Exception as e:
    print('$f', 'FAILED', e)
PY
done
# 3. 3-D INS-DIM on one GPU, FP64 and FP32 storage (N = 48: 2.9 M DOFs)
timeout 400 python bench.py --workload ins3d_dim --N 48 --steps 2 --warmup 1 --no-cpu > $O/bench3d_48.json 2> $O/bench3d_48.err
timeout 400 python bench.py --workload ins3d_dim --N 48 --steps 2 --warmup 1 --no-cpu --precond-storage fp32 > $O/bench3d_48_fp32.json 2> $O/bench3d_48_fp32.err
tail -c 600 $O/bench3d_48.json; echo; tail -c 600 $O/bench3d_48_fp32.json; echo
# 3b. N = 64 (6.8 M DOFs, nnz 1.42e9): failed in round 1 with an illegal address — int32 midpoint overflow in the
#     patch-position bisection, fixed since; bounded by its own timeout
timeout 600 python bench.py --workload ins3d_dim --N 64 --steps 2 --warmup 1 --no-cpu > $O/bench3d_64.json 2> $O/bench3d_64.err
tail -c 600 $O/bench3d_64.json; echo; tail -3 $O/bench3d_64.err
# 4. launch list + one full capture of the FP32 smoother kernel (a number printed under ncu is never a bench value)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_fp32.csv \
    python bench.py --N 128 --steps 1 --warmup 1 --no-cpu --precond-storage fp32 > $O/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_patch_apply_f32 -s 20 -c 6 -o $O/patch_apply_f32 -f \
    python bench.py --N 128 --steps 1 --warmup 1 --no-cpu --precond-storage fp32 > $O/ncu_full.log 2>&1
ls -la $O
