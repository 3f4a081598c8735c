cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2h_tests.log
tail -8 gpurun_out/r2h_tests.log
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-secondary --parity-sample 100 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "rc=$?" >> gpurun_out/r2h_bench.err
tail -2 gpurun_out/r2h_bench.err | cut -c1-300
timeout 300 python tools/time_decode.py > gpurun_out/r2h_k1.log 2>&1; cat gpurun_out/r2h_k1.log
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --parity-sample 0"
timeout 600 $CMD > gpurun_out/r2h_plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:SearchKernel -s 4 -c 1 -o gpurun_out/r2h_two $CMD > gpurun_out/r2h_ncu_f.log 2>&1
