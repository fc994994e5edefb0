#!/bin/bash
# round 2, call 3O (8 GPUs, second session): multi-GPU check on 8 ranks, copy probe at 1/2/4/8 GPUs, the bench line through torchrun
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_multigpu_gpu.py -m gpu -q > gpurun_out/r3o_pytest.log 2>&1; tail -4 gpurun_out/r3o_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r3o_bench8.log 2> gpurun_out/r3o_bench8.err; echo "bench8 rc=$?"; tail -3 gpurun_out/r3o_bench8.err
tail -1 gpurun_out/r3o_bench8.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('n_gpus', d['n_gpus'], 'value %.4e' % d['value'], 'frac %.3f' % d['roofline']['frac'], 'e2e %.4e' % d['e2e']['value'])
print(json.dumps(d['e2e'].get('copy_ceiling'), indent=1))
print(json.dumps(d['extra'].get('stats_check'), indent=1))
print({k: (v.get('env_steps_per_s') or v.get('ms')) for k, v in d['extra'].items() if isinstance(v, dict)})"
