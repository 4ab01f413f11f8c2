#!/bin/bash
# Run on the GPU box (via gpurun): GPU parity tests, one pytest process per file so a hung kernel
# cannot take the other files down; logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for f in ${@:-test_gpu_lap test_gpu_cost test_gpu_path}; do
  timeout 600 python -m pytest tests/$f.py -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider \
      > gpurun_out/$f.log 2>&1
  echo "== $f exit $?"; tail -n 25 gpurun_out/$f.log
done
