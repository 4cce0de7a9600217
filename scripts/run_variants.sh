#!/bin/bash
# usage (on the GPU box): scripts/run_variants.sh OUT.log workload steps variant...   ("default" = the shipped library)
# Runs bench.py's device-resident leg once per tuning variant (libmassiv_b200_<variant>.so, built with
# MB_VARIANT=<variant> MB_NVCC_EXTRA=... python -m massiv_b200.build) so that variants are compared on one box.
out=$1; wl=$2; steps=$3; shift 3
for v in "$@"; do
  lib=massiv_b200/libmassiv_b200_$v.so
  [ "$v" = default ] && lib=massiv_b200/libmassiv_b200.so
  echo "== $v" >> $out
  MASSIV_B200_LIB=$PWD/$lib python bench.py --workload $wl --no-e2e --no-cpu --no-verify --configs none --steps $steps --warmup 5 2>&1 \
    | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.rstrip()[-300:]); continue
    print(json.dumps({'value': round(d['value'], 1), 'ms': round(d['ms_per_step'], 4), 'kernel': d['config']['kernel'], 'frac': round(d['roofline']['frac'], 3), 'clocks': d['clocks']}))
" >> $out
done
