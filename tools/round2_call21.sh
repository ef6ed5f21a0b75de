#!/bin/bash
# 1-GPU call: Cannon tests with ranks sharing GPU 0 (gloo; includes the distributed-input case), DRAM traffic per multiply of both stack orders
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cannon.py -m gpu -q 2>&1 | tail -15 | tee gpurun_out/call21_tests.log
timeout 600 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
  --csv --log-file gpurun_out/dram_traffic.csv python tools/dram_traffic.py 2>&1 | tail -3
python tools/dram_traffic.py --digest gpurun_out/dram_traffic.csv | tee gpurun_out/dram_traffic.json
