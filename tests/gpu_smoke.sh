#!/bin/bash
# first-contact GPU run: tests, then a tiny bench
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -40
