#!/bin/bash
cd /root/repo
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2d_pytest.log
timeout 900 python tools/ab_bench.py --scale 0.5 "kernel=3" "kernel=4" "kernel=4,threads=768" "kernel=4,threads=512" "kernel=4,l2_hints=1" "kernel=3" > gpurun_out/r2d_ab.txt 2>&1
DACH_LIB=/root/repo/tools/alt/lib_q6.so timeout 900 python tools/ab_bench.py --scale 0.5 "kernel=3" "kernel=4" "kernel=4,threads=768" >> gpurun_out/r2d_ab.txt 2>&1
cat gpurun_out/r2d_ab.txt
timeout 600 python tools/ab_bench.py --scale 0.5 --mode find "kernel=3" "kernel=4" > gpurun_out/r2d_ab_find.txt 2>&1; cat gpurun_out/r2d_ab_find.txt
timeout 600 python tools/ab_bench.py --config C2 --scale 1 "kernel=3" "kernel=4" "kernel=2" > gpurun_out/r2d_ab_c2.txt 2>&1; cat gpurun_out/r2d_ab_c2.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scan_duo -s 1 -c 1 -f -o gpurun_out/prof_r2d_duo python bench.py --steps 2 --warmup 1 --scale 0.25 --no-e2e --no-cpu --option kernel=4 > gpurun_out/r2d_ncu.log 2>&1; echo "ncu rc=$?"
