bash scripts/gpu_check.sh test_gpu_metrics test_gpu_lap test_gpu_cost test_gpu_path
python scripts/gpu_metric_bench.py 10000 20000 > gpurun_out/metric_bench.log 2>&1; cat gpurun_out/metric_bench.log
