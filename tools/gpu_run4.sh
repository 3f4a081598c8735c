cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2d_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2d_tests.log
tail -15 gpurun_out/r2d_tests.log
# in-process group over two GPUs (NCCL between the devices of one process) vs the reference's results
G=tests/golden
head -c -0 $G/zipf2k/queries.txt | grep -v '^"' > /tmp/q_noph.txt
timeout 300 ./wiser_b200/wsr_replay -dirs=$G/zipf2k_p0,$G/zipf2k_p1 -devices=0,1 -query_path=/tmp/q_noph.txt -n_results=10 -batch_size=900 -dump=gpurun_out/r2d_replay2.txt > gpurun_out/r2d_replay2.log 2>&1; echo "replay rc=$?" >> gpurun_out/r2d_replay2.log
timeout 300 ./wiser_b200/wsr_replay -dirs=$G/zipf2k_p0,$G/zipf2k_p1 -devices=0 -query_path=/tmp/q_noph.txt -n_results=10 -batch_size=900 -dump=gpurun_out/r2d_replay1.txt > gpurun_out/r2d_replay1.log 2>&1
cmp gpurun_out/r2d_replay1.txt gpurun_out/r2d_replay2.txt && echo "replay 1-GPU and 2-GPU dumps identical" >> gpurun_out/r2d_replay2.log
tail -3 gpurun_out/r2d_replay2.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 3 --parity-sample 100 > gpurun_out/r2d_weak2.json 2> gpurun_out/r2d_weak2.err; echo "rc=$?" >> gpurun_out/r2d_weak2.err
tail -4 gpurun_out/r2d_weak2.err
timeout 900 $TR bench.py --gpus 2 --scaling strong --total-parts 4 --steps 10 --warmup 3 --parity-sample 50 > gpurun_out/r2d_strong4_n2.json 2> gpurun_out/r2d_strong4_n2.err; echo "rc=$?" >> gpurun_out/r2d_strong4_n2.err
tail -4 gpurun_out/r2d_strong4_n2.err
# single-GPU A/B of the filter density and K1
for v in "2 256" "4 256" "4 1024" "3 1024"; do set -- $v
  WSR_FILTER_PPW=$1 WSR_FILTER_MIN_DF=$2 timeout 600 python bench.py --steps 20 --warmup 3 --no-secondary --no-cpu-baseline --parity-sample 50 > gpurun_out/r2d_flt_$1_$2.json 2> gpurun_out/r2d_flt_$1_$2.err; echo "rc=$?" >> gpurun_out/r2d_flt_$1_$2.err
done
ls gpurun_out | grep r2d
