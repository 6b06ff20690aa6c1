"""Key ncu --set full metrics of every kernel in a .ncu-rep: python scratch/ncu_quick.py gpurun_out/x.ncu-rep"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
hh = rr[0]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.max.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "lts__t_sectors_op_red.sum", "lts__t_requests_op_red.sum" if False else "lts__t_sectors_op_atom.sum"]
for r in rr[2:]:
    print(r[hh.index("Kernel Name")][:60], r[hh.index("Grid Size")] if "Grid Size" in hh else "")
    for k in keys:
        if k in hh: print(f"   {k:75s} {r[hh.index(k)]} {rr[1][hh.index(k)]}")
    st = []
    for i, c in enumerate(hh):
        if c.startswith("smsp__pcsamp_warps_issue_stalled_") and not c.endswith("_not_issued"):
            try: st.append((float(r[i]), c.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError: pass
    ts = sum(v for v, _ in st) or 1
    print("   stalls:", ", ".join(f"{c} {100*v/ts:.0f}%" for v, c in sorted(st, reverse=True)[:6]))
