# Developer tool: full single-GPU validation on a B200 box (under gpurun): GPU tests, smoke, reference
# arm, default bench with trace, ncu launch list, ncu --set full of the dominant kernel, K1 timing.
# usage: gpurun -- 'TAG=<name> [STRONG1=1] bash tools/gpu_validate.sh'
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=${TAG:-run}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
tail -4 gpurun_out/${T}_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${T}_smoke.log; tail -2 gpurun_out/${T}_smoke.log
timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${T}_ref.json 2> gpurun_out/${T}_ref.err; echo "rc=$?" >> gpurun_out/${T}_ref.err
WSR_TRACE=1 timeout 1500 python bench.py --steps 20 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "rc=$?" >> gpurun_out/${T}_bench.err
grep "wsr trace" gpurun_out/${T}_bench.err | tail -2; tail -1 gpurun_out/${T}_bench.err | cut -c1-200
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --parity-sample 0"
timeout 600 $CMD > gpurun_out/${T}_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv $CMD > gpurun_out/${T}_ncu_l.log 2>&1
timeout 600 $CMD > gpurun_out/${T}_plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:SearchKernel -s 4 -c 1 -o gpurun_out/${T}_two $CMD > gpurun_out/${T}_ncu_f.log 2>&1
timeout 300 python tools/time_decode.py > gpurun_out/${T}_k1.log 2>&1; cat gpurun_out/${T}_k1.log | tail -3
if [ -n "$STRONG1" ]; then bash tools/gpu_run_n.sh 1; fi
