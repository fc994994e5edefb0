# host topology of the box + the e2e leg with and without the GPU-local CPU binding (bench.py --e2e-steps 3, extras off)
nvidia-smi topo -m 2>&1 | head -20
echo "nproc $(nproc)  nodes: $(ls -d /sys/devices/system/node/node* 2>/dev/null | wc -l)"
for n in /sys/devices/system/node/node*; do echo "$n $(cat $n/cpulist)"; done
python - <<'PY'
import os, torch
from cantorrl_b200.distributed import gpu_local_cpus
p = torch.cuda.get_device_properties(0)
bus = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
print("gpu0", bus, "numa_node", open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip() if os.path.exists(f"/sys/bus/pci/devices/{bus}/numa_node") else "n/a",
      "local cpus", len(gpu_local_cpus(bus)), "allowed", len(os.sched_getaffinity(0)))
PY
Q="--steps 5 --warmup 3 --e2e-steps 3 --rollout-steps 0 --mlp-rollout-steps 0 --lstm-rollout-steps 0 --book-strikes 0 --rbergomi-paths 0 --no-cpu-baseline"
for i in 1 2; do
CANTOR_NO_NUMA_BIND=1 timeout 200 python bench.py $Q 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('unbound', d['e2e']['value'], d['config']['host_numa_bind'])"
timeout 200 python bench.py $Q 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bound  ', d['e2e']['value'], d['config']['host_numa_bind'])"
done
