#!/bin/bash
# round 2, call 3N (2 GPUs, second session): the multi-GPU tests that skip on one GPU, then the bench line through torchrun, the copy probe on both GPUs
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_multigpu_gpu.py -m gpu -q > gpurun_out/r3n_pytest.log 2>&1; tail -12 gpurun_out/r3n_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r3n_bench2.log 2> gpurun_out/r3n_bench2.err; echo "bench2 rc=$?"; tail -5 gpurun_out/r3n_bench2.err
tail -1 gpurun_out/r3n_bench2.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('n_gpus', d['n_gpus'], 'value %.4e' % d['value'], 'frac %.3f' % d['roofline']['frac'], 'e2e %.4e' % d['e2e']['value'])
print(json.dumps(d['e2e'].get('copy_ceiling'), indent=1))
print(json.dumps(d['extra'].get('stats_check'), indent=1))
print({k: (v.get('env_steps_per_s') or v.get('ms')) for k, v in d['extra'].items() if isinstance(v, dict)})"
