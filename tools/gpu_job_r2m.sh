#!/bin/bash
# round 2, call M: persistent kernel, prefetch two steps ahead: parity tests on the variant build, then timing A/B
mkdir -p gpurun_out
CANTOR_HEDGE_LIB=build/variants/many_pf2/libcantor_hedge.so timeout 600 python -m pytest tests/test_step_modes_gpu.py -m gpu -q -k "step_many" 2>&1 | tail -3
MIN="--e2e-steps 0 --no-cpu-baseline --rollout-steps 0 --mlp-rollout-steps 0 --lstm-rollout-steps 0 --book-strikes 0 --rbergomi-paths 0 --l2free-envs 0 --no-forms"
one() { CANTOR_HEDGE_LIB=$1 timeout 300 python bench.py $MIN --steps 20 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$2 value=%.4e sweep_us=%.1f frac=%.3f' % (d['value'], d['roofline']['launch_us'], d['roofline']['frac']))"; }
for i in 1 2; do
  one cantorrl_b200/csrc/libcantor_hedge.so shipped_pf1
  one build/variants/many_pf2/libcantor_hedge.so many_pf2
  one build/variants/many_pf2_b9/libcantor_hedge.so many_pf2_b9
  one build/variants/many_pf2_b8/libcantor_hedge.so many_pf2_b8
done
