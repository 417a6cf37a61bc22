"""Summarise an .ncu-rep (read on the CPU box): key metrics of the profiled launches + the hottest
source lines by stall samples.  Usage: python scripts/ncu_summary.py gpurun_out/prof_search.ncu-rep"""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_uniform.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_drain_per_issue_active.ratio", "smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_misc_per_issue_active.ratio"]
for r in rows[2:]:
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            if r[i] not in ("", "-nan", "nan"):
                print(f"{w} = {r[i]} {units[i]}")
    print("---")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"], capture_output=True, text=True).stdout
try:
    srows = list(csv.reader(io.StringIO(src)))
    # find header row containing 'Source'
    hi = next(i for i, r in enumerate(srows) if "Source" in r and any("Sampling" in c for c in r))
    h = srows[hi]
    si = h.index("Source")
    samp = next(i for i, c in enumerate(h) if c.startswith("# Samples") or c.startswith("Warp Stall Sampling (All"))
    top = []
    for r in srows[hi + 1:]:
        try:
            top.append((float(r[samp] or 0), r[0], r[si].strip()[:110]))
        except Exception:
            pass
    tot = sum(t[0] for t in top) or 1
    print("hottest source lines by stall samples (share of all samples):")
    for s, ln, text in sorted(top, reverse=True)[:25]:
        print(f"  {100 * s / tot:5.1f}%  line {ln}: {text}")
except Exception as e:
    print("source page not parsed:", e)
