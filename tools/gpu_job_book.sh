CMD="python tools/bench_book.py --reps 1"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:book_f32_kernel -s 1 -c 1 -f -o gpurun_out/prof_book_r1 $CMD > gpurun_out/ncu_book.log 2>&1; echo "ncu rc=$?"
python tools/ncu_lines.py gpurun_out/prof_book_r1.ncu-rep cantorrl_b200/csrc/book_f32.o book_f32_kernel 60 > gpurun_out/book_lines.txt 2>&1; tail -1 gpurun_out/book_lines.txt
python tools/ncu_summary.py gpurun_out/prof_book_r1.ncu-rep book_f32_kernel 0 > gpurun_out/book_summary.txt 2>&1
timeout 100 python tools/bench_book.py | tail -1 | cut -c1-400
