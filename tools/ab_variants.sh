#!/bin/bash
# Developer tool: A/B of library variants on one GPU box (same corpus, same log). Build variants as
# _var/libwsr_<name>.so (e.g. with -D switches), then: gpurun -- bash tools/ab_variants.sh base <name>[:ppw[:workload]] ...
cp wiser_b200/libwsr.so /tmp/libwsr_base.so
for spec in "$@"; do
  IFS=: read v ppw wl <<< "$spec"
  ppw=${ppw:-4}; wl=${wl:-two_term}
  if [ $v = base ]; then cp /tmp/libwsr_base.so wiser_b200/libwsr.so; else cp _var/libwsr_$v.so wiser_b200/libwsr.so; fi
  echo "== $spec"
  WSR_FILTER_PPW=$ppw python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline --no-secondary --parity-sample 50 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        j = json.loads(l); print(j['ms_per_step'], j['value'], j['roofline']['frac'], j['e2e']['ms_per_step'])"
done
cp /tmp/libwsr_base.so wiser_b200/libwsr.so
