#!/bin/bash
# round 2, call 3B: tcgen05.mma cost vs N, SS and TS forms; clock sampler check on a minimal bench run
mkdir -p gpurun_out
timeout 120 python tools/umma_probe.py --ts 0 > gpurun_out/r3b_umma_ss.txt 2>&1; echo "ss rc=$?"; grep -v "^\[" gpurun_out/r3b_umma_ss.txt | head -32
timeout 120 python tools/umma_probe.py --ts 1 > gpurun_out/r3b_umma_ts.txt 2>&1; echo "ts rc=$?"; grep -v "^\[" gpurun_out/r3b_umma_ts.txt | head -32


