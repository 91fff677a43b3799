#!/bin/bash
cd /root/repo
timeout 300 python -m pytest tests -m gpu -x -q -k "across_devices or shard_group" > gpurun_out/r2n2_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2n2_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > gpurun_out/r2n2_bench.json 2> gpurun_out/r2n2_bench.err; echo "bench n2 rc=$?"; tail -3 gpurun_out/r2n2_bench.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/r2n2_bench.json')); print(d['value'], d['ms_per_step'], d['config']['peer_push_ms_by_rank'], d['cpu_baseline']['parity']); [print(r,t) for r,t in enumerate(d['config']['timeline_last4_by_rank'])]"
