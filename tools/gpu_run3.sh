cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2c_tests.log
tail -15 gpurun_out/r2c_tests.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-secondary --no-cpu-baseline --parity-sample 50 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "rc=$?" >> gpurun_out/r2c_bench.err
tail -2 gpurun_out/r2c_bench.err
timeout 900 python bench.py --scaling strong --total-parts 4 --steps 10 --warmup 3 --parity-sample 50 > gpurun_out/r2c_strong4.json 2> gpurun_out/r2c_strong4.err; echo "rc=$?" >> gpurun_out/r2c_strong4.err
tail -5 gpurun_out/r2c_strong4.err
