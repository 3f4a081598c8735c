# usage: bash tools/gpu_run_n.sh N   (under gpurun --gpus N)
N=$1
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/r2n${N}_gpus.txt
if [ "$N" = "1" ]; then
  timeout 1500 python bench.py --scaling strong --total-parts 8 --steps 10 --warmup 3 --parity-sample 50 > gpurun_out/r2n1_strong8.json 2> gpurun_out/r2n1_strong8.err; echo "rc=$?" >> gpurun_out/r2n1_strong8.err
  grep -v "^\[W\|Setting OMP" gpurun_out/r2n1_strong8.err | tail -4 | cut -c1-300
  exit 0
fi
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 3 --parity-sample 100 > gpurun_out/r2n${N}_weak.json 2> gpurun_out/r2n${N}_weak.err; echo "rc=$?" >> gpurun_out/r2n${N}_weak.err
grep -v "^\[W\|Setting OMP" gpurun_out/r2n${N}_weak.err | tail -4 | cut -c1-300
[ -n "$WEAK_ONLY" ] && exit 0
timeout 900 $TR bench.py --gpus $N --scaling strong --total-parts 8 --steps 10 --warmup 3 --parity-sample 50 > gpurun_out/r2n${N}_strong8.json 2> gpurun_out/r2n${N}_strong8.err; echo "rc=$?" >> gpurun_out/r2n${N}_strong8.err
grep -v "^\[W\|Setting OMP" gpurun_out/r2n${N}_strong8.err | tail -4 | cut -c1-300
if [ "$N" = "2" ]; then
  D=/tmp/wsr_bench
  DIRS=$D/c_d5000000_v5000000_mu5.34_s1_p0of8,$D/c_d5000000_v5000000_mu5.34_s1_p1of8
  timeout 600 ./wiser_b200/wsr_replay -dirs=$DIRS -devices=0,1 -query_path=$D/c_d5000000_v5000000_mu5.34_s1_p0of8/q_two_term_n100000_h10000_s1.txt -n_results=10 -batch_size=100000 -repeat=8 > gpurun_out/r2n2_replay_inproc.log 2>&1; echo "rc=$?" >> gpurun_out/r2n2_replay_inproc.log
  tail -3 gpurun_out/r2n2_replay_inproc.log | cut -c1-500
fi
if [ "$N" = "8" ]; then
  # one process over 8 GPUs (in-process NCCL ranks) through the replay driver
  D=/tmp/wsr_bench
  DIRS=$(for p in 0 1 2 3 4 5 6 7; do printf "%s/c_d5000000_v5000000_mu5.34_s1_p%dof8," $D $p; done | sed 's/,$//')
  timeout 600 ./wiser_b200/wsr_replay -dirs=$DIRS -devices=0,1,2,3,4,5,6,7 -query_path=$D/c_d5000000_v5000000_mu5.34_s1_p0of8/q_two_term_n100000_h10000_s1.txt -n_results=10 -batch_size=100000 -repeat=6 > gpurun_out/r2n8_replay_inproc.log 2>&1; echo "rc=$?" >> gpurun_out/r2n8_replay_inproc.log
  tail -3 gpurun_out/r2n8_replay_inproc.log | cut -c1-400
fi
