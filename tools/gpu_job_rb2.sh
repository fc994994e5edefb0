# the FIRST price launch of tools/bench_rbergomi.py is its 64-path x 64-inner-path warm-up (one iteration per CTA): skip it (-s 1)
CMD="python tools/bench_rbergomi.py --paths 512 --steps 32"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rbergomi_price -s 1 -c 1 -f -o gpurun_out/prof_rbergomi_price_tc_r1 $CMD > gpurun_out/ncu_rb.log 2>&1; echo "ncu rc=$?"
python tools/ncu_lines.py gpurun_out/prof_rbergomi_price_tc_r1.ncu-rep cantorrl_b200/csrc/rbergomi.o rbergomi_price 60 > gpurun_out/rb_lines.txt 2>&1; tail -1 gpurun_out/rb_lines.txt
python tools/ncu_summary.py gpurun_out/prof_rbergomi_price_tc_r1.ncu-rep rbergomi_price 0 > gpurun_out/rb_summary.txt 2>&1
