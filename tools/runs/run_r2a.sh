#!/bin/bash
cd /root/repo
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"
timeout 600 python tools/ab_bench.py --scale 0.5 "kernel=3" "kernel=3,hot_entries=0" "kernel=3,hot_entries=4096" "kernel=3,hot_entries=8192" "kernel=2" "kernel=3,threads=768,ctas_per_sm=2" "kernel=3,threads=512,ctas_per_sm=2" "kernel=3,l2_persist=0" "kernel=3,threads=768" "kernel=3" > gpurun_out/r2a_ab.txt 2>&1
DACH_LIB=/root/repo/tools/alt/lib_generic.so timeout 600 python tools/ab_bench.py --scale 0.5 "kernel=3" "kernel=3,hot_entries=4096" "kernel=3,hot_entries=0" >> gpurun_out/r2a_ab.txt 2>&1
timeout 900 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scan_machine -s 1 -c 1 -f -o gpurun_out/prof_r2a python bench.py --steps 2 --warmup 1 --scale 0.25 --no-e2e --no-cpu > gpurun_out/r2a_ncu.log 2>&1; echo "ncu rc=$?"
cat gpurun_out/r2a_ab.txt; tail -3 gpurun_out/r2a_pytest.log; cat gpurun_out/r2a_bench.json | head -c 1500
