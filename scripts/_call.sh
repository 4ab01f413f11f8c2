bash scripts/gpu_check.sh test_gpu_lap
python scripts/gpu_lap_sweep3.py 10000 20000 1 1002,1003 CYB_LAP_TAIL_MODE=1 CYB_LAP_TAIL=8,16,32 > gpurun_out/sweep_10k.log 2>&1; cat gpurun_out/sweep_10k.log
python scripts/gpu_lap_sweep3.py 30000 6000 6 1004 CYB_LAP_TAIL_MODE=1 CYB_LAP_TAIL=8,32 > gpurun_out/sweep_cfg4.log 2>&1; cat gpurun_out/sweep_cfg4.log
python scripts/gpu_lap_sweep3.py 25000 20000 1 1005 CYB_LAP_TAIL_MODE=1 CYB_LAP_TAIL=8,32 > gpurun_out/sweep_25k.log 2>&1; cat gpurun_out/sweep_25k.log
