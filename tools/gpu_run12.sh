cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=${TAG:-r2p}
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/${T}_tests.txt
timeout 400 bash tools/ab_variants.sh $VARIANTS > gpurun_out/${T}_ab.txt 2>&1; cat gpurun_out/${T}_tests.txt gpurun_out/${T}_ab.txt
if [ -n "$NCU" ]; then
cp wiser_b200/libwsr.so /tmp/keep.so; [ "$NCU" != base ] && cp _var/libwsr_$NCU.so wiser_b200/libwsr.so
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --parity-sample 0"
timeout 600 $CMD > gpurun_out/${T}_plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:SearchKernel -s 4 -c 1 -o gpurun_out/${T}_two $CMD > gpurun_out/${T}_ncu_f.log 2>&1
cp /tmp/keep.so wiser_b200/libwsr.so
fi
