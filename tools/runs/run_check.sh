#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_check_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_check_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_check_bench.json 2> gpurun_out/r2_check_bench.err; echo "bench rc=$?"; head -c 250 gpurun_out/r2_check_bench.json; echo; tail -2 gpurun_out/r2_check_bench.err
