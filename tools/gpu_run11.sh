cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 500 bash tools/ab_variants.sh base nopf nopfp nopff > gpurun_out/r2m_ab.txt 2>&1; cat gpurun_out/r2m_ab.txt
cp wiser_b200/libwsr.so /tmp/keep.so; cp _var/libwsr_nopf.so wiser_b200/libwsr.so
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --parity-sample 0"
timeout 600 $CMD > gpurun_out/r2m_plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:SearchKernel -s 4 -c 1 -o gpurun_out/r2m_two $CMD > gpurun_out/r2m_ncu_f.log 2>&1
cp /tmp/keep.so wiser_b200/libwsr.so
ls -la gpurun_out/
