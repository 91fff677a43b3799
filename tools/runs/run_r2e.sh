#!/bin/bash
cd /root/repo
timeout 600 python tools/ab_bench.py --scale 0.5 "kernel=3" "kernel=4" "kernel=4,hot_entries=-1" "kernel=4,hot_entries=2048" "kernel=4,hot_entries=-1,threads=768" > gpurun_out/r2e_ab.txt 2>&1
DACH_LIB=/root/repo/tools/alt/lib_q6.so timeout 600 python tools/ab_bench.py --scale 0.5 "kernel=3" "kernel=4" "kernel=4,hot_entries=-1" "kernel=4,hot_entries=4096" "kernel=3,hot_entries=-1" >> gpurun_out/r2e_ab.txt 2>&1
DACH_LIB=/root/repo/tools/alt/lib_q5.so timeout 600 python tools/ab_bench.py --scale 0.5 "kernel=3" "kernel=4,hot_entries=-1" "kernel=4,hot_entries=6144" >> gpurun_out/r2e_ab.txt 2>&1
cat gpurun_out/r2e_ab.txt
timeout 1200 python -m pytest tests -m gpu -x -q -k "stream or sharding or group or jobs or options" > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2e_pytest.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize.py > gpurun_out/r2e_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/r2e_memcheck.log
