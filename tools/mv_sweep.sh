#!/bin/bash
# mat-vec timings: ring x kappa x log2(m) on one GPU (CUDA events, min of 5).  usage: tools/mv_sweep.sh "gl bb sp" "1 2 4 8" "17 20"
RINGS=${1:-"gl bb sp"}; KS=${2:-"1 2 4 8"}; LS=${3:-"17 20"}
for ring in $RINGS; do
  for kappa in $KS; do
    for l in $LS; do
      python tools/run_op.py --ring $ring --op matvec --kappa $kappa --log2n $l --reps 5 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('%s kappa=$kappa m=2^%d  %.4f ms  %.0f GB/s  frac %.3f' % (d['ring'], d['log2n'], d['ms_min'], d['GBps'], d['frac_hbm_6536']))"
    done
  done
done
