cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2a_tests.log
tail -5 gpurun_out/r2a_tests.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "rc=$?" >> gpurun_out/r2a_bench.err
tail -3 gpurun_out/r2a_bench.err
for r in 0 8 32; do
  WSR_MERGE_RATIO_X4=$r timeout 600 python bench.py --steps 20 --warmup 3 --no-secondary --no-cpu-baseline --parity-sample 50 > gpurun_out/r2a_bench_mr$r.json 2> gpurun_out/r2a_bench_mr$r.err; echo "rc=$?" >> gpurun_out/r2a_bench_mr$r.err
done
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --parity-sample 0"
timeout 600 $CMD > gpurun_out/r2a_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a_launches.csv $CMD > gpurun_out/r2a_ncu_l.log 2>&1
timeout 600 $CMD > gpurun_out/r2a_plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:SearchKernel -s 4 -c 1 -o gpurun_out/r2a_two $CMD > gpurun_out/r2a_ncu_f.log 2>&1
ls -la gpurun_out | tail -20
