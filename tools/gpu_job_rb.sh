timeout 300 python tools/bench_rbergomi.py --paths 2048 --no-tc 2>&1 | tail -1
CMD="python tools/bench_rbergomi.py --paths 512 --steps 32"
timeout 200 $CMD > gpurun_out/plain_rb.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rbergomi_p -s 2 -c 2 -f -o gpurun_out/prof_rbergomi_price_tc_r1 $CMD > gpurun_out/ncu7.log 2>&1
echo "ncu7 rc=$?"
