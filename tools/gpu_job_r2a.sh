#!/bin/bash
# round 2, call A: the step kernel where nothing survives in L2 (2^23 envs), variants, raw host-copy ceiling, DRAM bytes per launch
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/r2a_gpu.txt
python tools/host_copy_probe.py --gpus 1 > gpurun_out/r2a_probe.log 2>&1
for n in 1048576 4194304 8388608; do
  timeout 300 python tools/bench_step_l2free.py --envs $n >> gpurun_out/r2a_step.log 2>&1
done
timeout 300 python tools/bench_step_l2free.py --precision fp64 >> gpurun_out/r2a_step.log 2>&1
for v in t256 noevict prefetch mb12; do
  CANTOR_HEDGE_LIB=build/variants/$v/libcantor_hedge.so timeout 300 python tools/bench_step_l2free.py >> gpurun_out/r2a_step.log 2>&1
done
timeout 600 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct \
  -k regex:hedge_step --launch-skip 300 --launch-count 60 --csv --log-file gpurun_out/r2a_step_2p23_dram.csv \
  python tools/bench_step_l2free.py --sweeps 1 > gpurun_out/r2a_ncu.log 2>&1
tail -n 12 gpurun_out/r2a_step.log; cat gpurun_out/r2a_probe.log
