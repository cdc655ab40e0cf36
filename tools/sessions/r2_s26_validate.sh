#!/bin/bash
# round-2 last validation of the tree as committed: build() (no-op when the .so is current) + smoke(), then the whole GPU suite
mkdir -p gpurun_out/s26
O=gpurun_out/s26
rm -f profiles/parity_gpu_latest.txt gpurun_out/parity_gpu_latest.txt
timeout 600 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 300 --timeout-method thread 2>&1 | tail -8 > $O/pytest_gpu.log
tail -4 $O/smoke.log; tail -4 $O/pytest_gpu.log
