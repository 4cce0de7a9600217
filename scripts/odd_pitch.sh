#!/bin/bash
# usage (on the GPU box): scripts/odd_pitch.sh OUT.log variant...   — the odd-pitch sizes of VERDICT r01 item 3
out=$1; shift
for v in "$@"; do
  lib=massiv_b200/libmassiv_b200_$v.so
  [ "$v" = default ] && lib=massiv_b200/libmassiv_b200.so
  for spec in "avg3x3 16383 16383" "sobel 32767 32767" "gauss7sep 32767 32767" "heat3d 1535 1535 1535"; do
    set -- $spec; wl=$1; shift
    echo "== $v $wl $*" >> $out
    MASSIV_B200_LIB=$PWD/$lib python bench.py --workload $wl --grid $* --no-e2e --no-cpu --no-verify --configs none --steps 10 --warmup 3 2>&1 \
      | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.rstrip()[-300:]); continue
    print(json.dumps({'value': round(d['value'], 1), 'ms': round(d['ms_per_step'], 4), 'kernel': d['config']['kernel'], 'frac': round(d['roofline']['frac'], 3)}))
" >> $out
  done
done
