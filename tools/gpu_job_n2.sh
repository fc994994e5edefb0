# two-GPU check: the multi-GPU tests that skip on one GPU, then the bench line through torchrun
timeout 300 python -m pytest tests/test_multigpu_gpu.py -m gpu -q 2>&1 | tail -2
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench2.log 2> gpurun_out/bench2.err; echo "bench2 rc=$?"
tail -1 gpurun_out/bench2.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('n_gpus', d['n_gpus'], 'value', d['value'], 'e2e', d['e2e']['value'], d['config'].get('host_numa_bind')); print({k:(v.get('env_steps_per_s') or v.get('ms')) for k,v in d['extra'].items()})"
