"""Brief per-kernel summary of an ncu --set full report: time, instructions, IPC, occupancy, pipe utilisation, stall mix.
    python tools/ncu_brief.py REPORT.ncu-rep"""
import csv, io, subprocess, sys
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
ix = {n: i for i, n in enumerate(hdr)}
K = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
     "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
     "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
     "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
     "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active",
     "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
     "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
     "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
     "l1tex__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    print("====", r[ix["Kernel Name"]][:70], r[ix["Grid Size"]], r[ix["Block Size"]])
    for k in K:
        if k in ix:
            print(f"   {k:75s} {r[ix[k]]} {units[ix[k]]}")
    st = [(n, r[i]) for n, i in ix.items() if n.startswith("smsp__average_warp") and "issue_stalled" in n and n.endswith("_per_issue_active.ratio")]
    st = sorted(((n.split("issue_stalled_")[1].split("_per")[0], float(v)) for n, v in st if v), key=lambda kv: -kv[1])[:8]
    print("   stalls (warp-cycles per issue):", ", ".join(f"{n}={v:.2f}" for n, v in st))
