#!/bin/bash
# round 2, call U: does the persistent step_many kernel lose time in its last partial wave?  whole-wave env counts vs 2^20
mkdir -p gpurun_out
for n in 1048576 947200 757760 1136640 1894400 2097152; do
  echo "--- envs $n"; timeout 300 python tools/bench_modes.py --mode many --envs $n --sweeps 6 | tail -1
done > gpurun_out/r2u_waves.log 2>&1
cat gpurun_out/r2u_waves.log
