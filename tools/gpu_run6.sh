cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2f_tests.log
tail -12 gpurun_out/r2f_tests.log
timeout 1200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "rc=$?" >> gpurun_out/r2f_bench.err
tail -3 gpurun_out/r2f_bench.err | cut -c1-400
timeout 300 python tools/time_decode.py > gpurun_out/r2f_k1.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:DecodeAllKernel -s 2 -c 1 -o gpurun_out/r2f_k1 python tools/time_decode.py > gpurun_out/r2f_k1_ncu.log 2>&1
cat gpurun_out/r2f_k1.log
