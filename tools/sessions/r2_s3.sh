#!/bin/bash
# round-2 GPU session 3: instruction-level profile of gemm_pair_kernel on the epilogue-bound token GEMMs
mkdir -p gpurun_out/s3
O=gpurun_out/s3
MDB_GEMM_VARIANT=3 timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none --cache-control none \
  -k regex:gemm_pair -o $O/prof_pair python tools/bench_gemm.py --profile --warm --only tok16800 > $O/ncu.log 2>&1
tail -5 $O/ncu.log
