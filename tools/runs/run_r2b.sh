#!/bin/bash
cd /root/repo
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2b_pytest.log
timeout 900 python tools/ab_bench.py --scale 0.5 "kernel=3" "l2_hints=0" "hot_entries=0" "hot_entries=0,l2_hints=0" "hot_entries=2048" "hot_entries=6144" "hot_entries=8192" "kernel=2" "kernel=2,l2_hints=0" "threads=768" "threads=768,hot_entries=6144" "kernel=3" > gpurun_out/r2b_ab.txt 2>&1
cat gpurun_out/r2b_ab.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scan_machine -s 1 -c 1 -f -o gpurun_out/prof_r2b python bench.py --steps 2 --warmup 1 --scale 0.25 --no-e2e --no-cpu > gpurun_out/r2b_ncu.log 2>&1; echo "ncu rc=$?"
