cat > /tmp/simrun.py <<'PY'
import sys; sys.path.insert(0, ".")
import torch
from cantorrl_b200 import sim
for _ in range(2):
    b = sim.generate_paths_and_options(1 << 20, model="gbm", n_steps=252)
torch.cuda.synchronize()
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sim_paths_kernel -s 1 -c 1 -f -o gpurun_out/prof_sim_r1 python /tmp/simrun.py > gpurun_out/ncu_sim.log 2>&1; echo "ncu rc=$?"
python tools/ncu_lines.py gpurun_out/prof_sim_r1.ncu-rep cantorrl_b200/csrc/path_sim.o sim_paths_kernel 60 > gpurun_out/sim_lines.txt 2>&1; tail -2 gpurun_out/sim_lines.txt
