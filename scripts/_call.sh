bash scripts/gpu_check.sh test_gpu_lap test_gpu_path
bash scripts/gpu_bench_all.sh cfg2 cfg4 chunk25k
python -c "
import json
for w in ('cfg2','cfg4','chunk25k'):
    d=json.load(open('gpurun_out/bench_%s.json'%w)); print(w, d['roofline_row_scan'])"
