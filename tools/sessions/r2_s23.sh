#!/bin/bash
# round-2 GPU session 23: tile width / split count of the split-K kernel on the 4x7-level convolutions (weight-streaming bound)
mkdir -p gpurun_out/s23
O=gpurun_out/s23
for bn in 0 128 160 256; do
  for sp in 0 6 9 12; do
    echo "## bn=$bn splits=$sp" >> $O/conv4x7.log
    MDB_GEMM_VARIANT=2 timeout 100 python tools/bench_gemm.py --warm --inner 10 --reps 5 --only conv4x7 --bn $bn --splits $sp >> $O/conv4x7.log 2>&1
  done
done
for v in 2 3; do for bn in 0 64 128 256; do
  echo "## variant=$v bn=$bn" >> $O/tok336.log
  MDB_GEMM_VARIANT=$v timeout 100 python tools/bench_gemm.py --warm --inner 10 --reps 5 --only tok336 --bn $bn >> $O/tok336.log 2>&1
done; done
for v in 2 3; do for bn in 0 128 256; do
  echo "## variant=$v bn=$bn" >> $O/conv7x13.log
  MDB_GEMM_VARIANT=$v timeout 100 python tools/bench_gemm.py --warm --inner 10 --reps 5 --only conv7x13 --bn $bn >> $O/conv7x13.log 2>&1
done; done
grep -v "^# variant" $O/conv4x7.log; grep -v "^# variant" $O/tok336.log; grep -v "^# variant" $O/conv7x13.log
