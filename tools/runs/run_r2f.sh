#!/bin/bash
cd /root/repo
timeout 600 python tools/ab_bench.py --scale 0.5 "kernel=3" "smem_pad_kib=64" "smem_pad_kib=100" "smem_pad_kib=140" "threads=768" "threads=768,smem_pad_kib=160" "kernel=3" > gpurun_out/r2f_ab.txt 2>&1; cat gpurun_out/r2f_ab.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scan_machine -s 1 -c 1 -f -o gpurun_out/prof_r2f_std3 python bench.py --steps 2 --warmup 1 --scale 0.25 --no-e2e --no-cpu > gpurun_out/r2f_ncu1.log 2>&1; echo "ncu std3 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scan_machine -s 1 -c 1 -f -o gpurun_out/prof_r2f_cw python bench.py --config C4 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2f_ncu2.log 2>&1; echo "ncu cw rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scan_machine -s 1 -c 1 -f -o gpurun_out/prof_r2f_lm python bench.py --config C4-bw --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2f_ncu3.log 2>&1; echo "ncu lm rc=$?"
timeout 600 python bench.py --config C4-bw --steps 6 --no-e2e > gpurun_out/r2f_bench_C4bw.json 2> gpurun_out/r2f_bench_C4bw.err; echo "bench C4-bw rc=$?"; head -c 300 gpurun_out/r2f_bench_C4bw.json; echo
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize.py > gpurun_out/r2f_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r2f_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 1 python tools/sanitize.py > gpurun_out/r2f_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r2f_racecheck.log
