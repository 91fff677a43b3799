#!/bin/bash
cd /root/repo
nvidia-smi topo -m > gpurun_out/r2_topo_n2.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q -k "across_devices or shard_group" > gpurun_out/r2n2_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2n2_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2n2_bench.json 2> gpurun_out/r2n2_bench.err; echo "bench n2 rc=$?"; tail -5 gpurun_out/r2n2_bench.err; head -c 1500 gpurun_out/r2n2_bench.json; echo
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 10 --warmup 3 --no-overlap --no-e2e --no-cpu > gpurun_out/r2n2_bench_noov.json 2> gpurun_out/r2n2_bench_noov.err; echo "bench n2 no-overlap rc=$?"; head -c 400 gpurun_out/r2n2_bench_noov.json; echo
timeout 900 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2n2_bench_n1.json 2> gpurun_out/r2n2_bench_n1.err; echo "bench n1 rc=$?"; head -c 400 gpurun_out/r2n2_bench_n1.json; echo
