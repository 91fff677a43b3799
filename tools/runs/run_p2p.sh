#!/bin/bash
cd /root/repo
timeout 300 python tools/p2p_probe.py 1024 > gpurun_out/r2_p2p_probe.txt 2>&1; cat gpurun_out/r2_p2p_probe.txt
