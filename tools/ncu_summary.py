"""Summarise an .ncu-rep: headline metrics per launch + top stall lines of the source page."""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "smsp__cycles_active.avg",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print(d.get("Kernel Name", "")[:80])
    for k in keys:
        for h in hdr:
            if h == k:
                print(f"   {k} = {d[h]} {units[hdr.index(h)]}")
    # stall breakdown (warp-state sampling)
    st = [(h, float(d[h])) for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and d[h] not in ("", "n/a")]
    st.sort(key=lambda x: -x[1])
    for h, v in st[:8]:
        print(f"   stall {h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')}: {v:.2f}")
    if "--all" not in sys.argv:
        break
if len(sys.argv) > 2 and "--all" not in sys.argv:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    print(src[:3000])
