#!/bin/bash
# Run on the GPU box: bench line + ncu launch list + ncu full captures of the two dominant kernels.
mkdir -p gpurun_out
W=${1:-cfg2}
python bench.py --workload $W --steps 5 --warmup 3 > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err
echo "bench exit $?"; cat gpurun_out/bench_$W.json; tail -3 gpurun_out/bench_$W.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$W.csv \
    python bench.py --workload $W --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_$W.log 2>&1
echo "ncu launches exit $?"; tail -2 gpurun_out/ncu_launch_$W.log
ncu --set full --clock-control none --import-source on -k regex:lap_auction -c 1 -f -o gpurun_out/prof_lap_$W \
    python bench.py --workload $W --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_lap_$W.log 2>&1
echo "ncu lap exit $?"
ncu --set full --clock-control none --import-source on -k regex:cost_gemm -c 1 -f -o gpurun_out/prof_gemm_$W \
    python bench.py --workload $W --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_gemm_$W.log 2>&1
echo "ncu gemm exit $?"
ls -la gpurun_out
