#!/bin/bash
cd /root/repo
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2c_pytest.log
timeout 900 python tools/ab_bench.py --scale 0.5 "l2_hints=1,hot_entries=0" "l2_hints=2,hot_entries=0" "l2_hints=2,hot_entries=2048" "l2_hints=2,hot_entries=4096" "l2_hints=0,hot_entries=0" "l2_hints=2,hot_entries=0,kernel=2" > gpurun_out/r2c_ab.txt 2>&1
DACH_HOT_SLOTS=0 timeout 900 python tools/ab_bench.py --scale 0.5 --tag norelayout "l2_hints=1,hot_entries=0" "l2_hints=2,hot_entries=0" "l2_hints=0,hot_entries=0" "l2_hints=0,hot_entries=0,kernel=2" >> gpurun_out/r2c_ab.txt 2>&1
cat gpurun_out/r2c_ab.txt
timeout 900 python bench.py > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2c_bench.err; head -c 2500 gpurun_out/r2c_bench.json
for c in C3-find C2 C4; do timeout 900 python bench.py --config $c --steps 6 > gpurun_out/r2c_bench_$c.json 2> gpurun_out/r2c_bench_$c.err; echo "bench $c rc=$?"; tail -2 gpurun_out/r2c_bench_$c.err; head -c 600 gpurun_out/r2c_bench_$c.json; echo; done
timeout 900 python bench.py --config C5 --scale 0.2 --steps 4 > gpurun_out/r2c_bench_C5s.json 2> gpurun_out/r2c_bench_C5s.err; echo "bench C5 rc=$?"; tail -2 gpurun_out/r2c_bench_C5s.err; head -c 600 gpurun_out/r2c_bench_C5s.json
