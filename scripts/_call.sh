bash scripts/gpu_check.sh test_gpu_metrics test_gpu_lap test_gpu_cost test_gpu_path
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
bash scripts/gpu_profile.sh cfg2 2>&1 | grep -v "^-rw\|^drwx\|^total"
bash scripts/gpu_bench_all.sh cfg3 cfg4 chunk25k
for M in Spearman_correlation Euclidean; do python bench.py --workload cfg2 --steps 3 --warmup 3 --no-cpu-baseline --distance-metric $M > gpurun_out/bench_cfg2_$M.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_cfg2_$M.json')); print('$M', d['ms_per_step'], d['lap_ms'], d['cost_build_ms'], d['certificate'])"; done
