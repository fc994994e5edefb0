# rough-Bergomi pricer: resident CTAs per SM (register cap) x second-stream draws issued under the MMA (CANTOR_RB_W2_EARLY)
for v in ${RB_VARIANTS:-"" rb5 rb4e rb5e}; do
  if [ "$v" != "default" ] && [ -n "$v" ]; then export CANTOR_HEDGE_LIB=build/variants/$v/libcantor_hedge.so; else unset CANTOR_HEDGE_LIB; fi
  echo "variant ${v:-default}: $(timeout 200 python tools/bench_rbergomi.py --paths 512 --steps 32 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['price_ms'], d['inner_path_steps_per_s'], d['mean_call'], d['mean_put'])")"
done
unset CANTOR_HEDGE_LIB
timeout 400 python -m pytest tests/test_rbergomi_gpu.py -m gpu -q 2>&1 | tail -1
