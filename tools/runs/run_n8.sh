#!/bin/bash
cd /root/repo
N=${1:-8}
shift
run() { tag=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus $N "$@" > gpurun_out/r2n${N}_$tag.json 2> gpurun_out/r2n${N}_$tag.err; echo "bench $tag rc=$?"; tail -2 gpurun_out/r2n${N}_$tag.err | cut -c1-300; head -c 330 gpurun_out/r2n${N}_$tag.json; echo; }
for what in "$@"; do
  case $what in
    C3) run C3 --steps 10 --warmup 3 ;;
    C3ne) run C3ne --steps 10 --warmup 3 --no-e2e ;;
    C5) run C5 --config C5 --steps 13 --warmup 3 --e2e-steps 1 ;;
    C3find) run C3find --config C3-find --steps 10 --warmup 3 --no-e2e ;;
  esac
done
