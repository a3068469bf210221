#!/bin/bash
# Final single-GPU session of round 2: GPU suite, both bench arms, ncu launch list and --set full captures
O=gpurun_out/r2_final
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -5 $O/smoke.log
( time timeout 600 python bench.py > $O/bench_1gpu.json 2> $O/bench_1gpu.err ) 2> $O/bench_1gpu.time; tail -c 300 $O/bench_1gpu.err; tail -3 $O/bench_1gpu.time
( time timeout 600 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err ) 2> $O/bench_reference.time; tail -c 300 $O/bench_reference.err; tail -3 $O/bench_reference.time
timeout 200 python tools/kern_bench3d.py 32 > $O/kern3d.log 2>&1; cat $O/kern3d.log
OCMP_PATCH_STORAGE=fp32 timeout 200 python tools/kern_bench.py 128 > $O/kern2d_fp32.log 2>&1; tail -9 $O/kern2d_fp32.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file $O/launches.csv \
    python bench.py --no-cpu --no-secondary --steps 1 --warmup 1 --profile-steps 0 > $O/ncu_launches.log 2>&1
gzip -f $O/launches.csv
for k in k_patch_apply_stream k_spmv_vec k_contract k_gather_add k_patch_invert k_mdot k_gs_finish; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 3 -o $O/ncu_$k -f python tools/kern_bench3d.py 32 > $O/ncu_$k.log 2>&1
done
ls -la $O
