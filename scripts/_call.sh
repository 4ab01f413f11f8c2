bash scripts/gpu_check.sh test_gpu_metrics test_gpu_lap test_gpu_cost test_gpu_path
bash scripts/gpu_profile.sh cfg2 2>&1 | tail -15
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg2_spearman.csv \
    python bench.py --workload cfg2 --steps 2 --warmup 1 --no-cpu-baseline --distance-metric Spearman_correlation > gpurun_out/ncu_launch_spearman.log 2>&1
bash scripts/gpu_bench_all.sh cfg3 cfg4 chunk25k
