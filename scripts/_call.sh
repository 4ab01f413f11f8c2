bash scripts/gpu_check.sh test_gpu_lap test_gpu_path
python scripts/gpu_lap_sweep3.py 30000 6000 6 1004,1008 CYB_LAP_PACKED=1 > gpurun_out/timing_cfg4.log 2>&1; cat gpurun_out/timing_cfg4.log
