"""Turn `ncu --set full` reports (gpurun_out/*.ncu-rep) into the small JSON summary committed under profiles/.
usage: ncu_summarise.py out.json name=report.ncu-rep ..."""
import csv, io, json, subprocess, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "smsp__cycles_active.avg", "sm__inst_executed_pipe_tensor.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum",
        "smsp__average_warp_latency_issue_stalled_barrier.pct", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]
out = {}
for arg in sys.argv[2:]:
    name, rep = arg.split("=")
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {"Kernel Name": vals[hdr.index("Kernel Name")]}
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            d[k] = f"{vals[i]} {units[i]}".strip()
    out[name] = d
json.dump(out, open(sys.argv[1], "w"), indent=1)
print(json.dumps(out, indent=1))
