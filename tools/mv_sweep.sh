#!/bin/bash
# mat-vec timings: ring x kappa x log2(m) on one GPU (CUDA events, min of 5)
for ring in gl bb sp; do
  for kappa in 1 2 4 8; do
    for l in 17 20; do
      python tools/run_op.py --ring $ring --op matvec --kappa $kappa --log2n $l --reps 5
    done
  done
done
