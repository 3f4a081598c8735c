cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2j_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2j_tests.log
tail -4 gpurun_out/r2j_tests.log
WSR_TRACE=1 timeout 900 python bench.py --steps 20 --warmup 3 --no-secondary --no-cpu-baseline --parity-sample 100 > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo "rc=$?" >> gpurun_out/r2j_bench.err
grep "wsr trace" gpurun_out/r2j_bench.err | tail -5
tail -1 gpurun_out/r2j_bench.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --parity-sample 0"
timeout 600 $CMD > gpurun_out/r2j_plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:SearchKernel -s 4 -c 1 -o gpurun_out/r2j_two $CMD > gpurun_out/r2j_ncu_f.log 2>&1
