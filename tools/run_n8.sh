#!/bin/bash
cd /root/repo
N=${1:-8}
nvidia-smi topo -m > gpurun_out/r2_topo_n$N.txt 2>&1
free -g | head -2 > gpurun_out/r2_mem_n$N.txt; nproc >> gpurun_out/r2_mem_n$N.txt; cat /sys/fs/cgroup/cpu.max >> gpurun_out/r2_mem_n$N.txt 2>&1
run() { tag=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus $N "$@" > gpurun_out/r2n${N}_$tag.json 2> gpurun_out/r2n${N}_$tag.err; echo "bench $tag rc=$?"; tail -2 gpurun_out/r2n${N}_$tag.err | cut -c1-300; head -c 330 gpurun_out/r2n${N}_$tag.json; echo; }
run C3 --steps 10 --warmup 3
run C3_noov --steps 10 --warmup 3 --no-overlap --no-e2e --no-cpu
run C5 --config C5 --steps 13 --warmup 3 --e2e-steps 1
run C3find --config C3-find --steps 10 --warmup 3 --no-e2e
