cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2i_tests.log
tail -8 gpurun_out/r2i_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2i_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2i_smoke.log; tail -2 gpurun_out/r2i_smoke.log
timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2i_ref.json 2> gpurun_out/r2i_ref.err; echo "rc=$?" >> gpurun_out/r2i_ref.err
timeout 1500 python bench.py --steps 20 --warmup 3 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "rc=$?" >> gpurun_out/r2i_bench.err
tail -2 gpurun_out/r2i_bench.err | cut -c1-200
WSR_NO_ZEROCOPY=1 timeout 900 python bench.py --steps 20 --warmup 3 --no-secondary --no-cpu-baseline --parity-sample 20 > gpurun_out/r2i_bench_nozc.json 2> gpurun_out/r2i_bench_nozc.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --parity-sample 0"
timeout 600 $CMD > gpurun_out/r2i_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2i_launches.csv $CMD > gpurun_out/r2i_ncu_l.log 2>&1
timeout 600 $CMD > gpurun_out/r2i_plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:SearchKernel -s 4 -c 1 -o gpurun_out/r2i_two $CMD > gpurun_out/r2i_ncu_f.log 2>&1
timeout 300 python tools/time_decode.py > gpurun_out/r2i_k1.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:DecodeAllKernel -s 2 -c 1 -o gpurun_out/r2i_k1 python tools/time_decode.py > gpurun_out/r2i_k1_ncu.log 2>&1
cat gpurun_out/r2i_k1.log
