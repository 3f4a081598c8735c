cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 3 --parity-sample 100 > gpurun_out/r2g_weak2.json 2> gpurun_out/r2g_weak2.err; echo "rc=$?" >> gpurun_out/r2g_weak2.err
grep -v "^\[W\|Setting OMP" gpurun_out/r2g_weak2.err | tail -6 | cut -c1-300
timeout 900 $TR bench.py --gpus 2 --scaling strong --total-parts 4 --steps 10 --warmup 3 --parity-sample 50 > gpurun_out/r2g_strong4_n2.json 2> gpurun_out/r2g_strong4_n2.err; echo "rc=$?" >> gpurun_out/r2g_strong4_n2.err
grep -v "^\[W\|Setting OMP" gpurun_out/r2g_strong4_n2.err | tail -6 | cut -c1-300
timeout 300 python tools/time_decode.py > gpurun_out/r2g_k1.log 2>&1; cat gpurun_out/r2g_k1.log
