#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_final_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_final_pytest.log
timeout 600 python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; echo "bench rc=$?"; head -c 300 gpurun_out/r2_final_bench.json; echo
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_final_bench_reference.json 2> gpurun_out/r2_final_bench_reference.err; echo "ref rc=$?"; head -c 200 gpurun_out/r2_final_bench_reference.json; echo
for c in C3-find C2 C4; do timeout 600 python bench.py --config $c > gpurun_out/r2_final_bench_$c.json 2> gpurun_out/r2_final_bench_$c.err; echo "bench $c rc=$?"; head -c 200 gpurun_out/r2_final_bench_$c.json; echo; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 80 --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2_final_launches.log 2>&1; echo "launch list rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_final_smoke.log
