bash scripts/gpu_check.sh test_gpu_lap
python scripts/gpu_lap_sweep3.py 10000 20000 1 1002,1003,1006,1007 CYB_LAP_HINTS=0,1 > gpurun_out/hints_10k.log 2>&1; cat gpurun_out/hints_10k.log
python scripts/gpu_lap_sweep3.py 30000 6000 6 1004 CYB_LAP_HINTS=0 > gpurun_out/hints_cfg4.log 2>&1; cat gpurun_out/hints_cfg4.log
