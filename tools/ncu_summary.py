"""Summarise an ncu report: key raw metrics + SASS regions (share of warp instructions, active threads per instruction)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
seg = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sass__thread_inst_executed_true_per_opcode",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio"]
for i, h in enumerate(hdr):
    if h in want:
        print(f"{h} [{rows[1][i]}]: {[r[i] for r in rows[2:]]}")
try:
    ti = hdr.index("sass__thread_inst_executed_true_per_opcode"); wi = hdr.index("smsp__inst_executed.sum")
    for r in rows[2:]:
        print("threads per warp instruction:", float(r[ti]) / float(r[wi]))
except ValueError:
    pass
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; blocks.append(cur); continue
    if r and r[0] == "Address":
        cur["hdr"] = r; continue
    if cur is not None and len(r) > 10:
        cur["rows"].append(r)
b = blocks[0]; ix = {k: i for i, k in enumerate(b["hdr"])}
W = lambda r: int(r[ix["Instructions Executed"]]); T = lambda r: int(r[ix["Thread Instructions Executed"]]); SM = lambda r: int(r[ix["# Samples"]])
tw, tt, ts = sum(map(W, b["rows"])), sum(map(T, b["rows"])), sum(map(SM, b["rows"]))
print(b["name"], "SASS instrs", len(b["rows"]), "warp inst", tw, "thread inst", tt, "thr/warp", tt / tw, "samples", ts)
for s in range(0, len(b["rows"]), seg):
    rr = b["rows"][s:s + seg]
    w, t, sm = sum(map(W, rr)), sum(map(T, rr)), sum(map(SM, rr))
    ops = {}
    for r in rr:
        src = r[ix["Source"]].split()
        op = (src[1] if src[0].startswith("@") else src[0]).split(".")[0]
        ops[op] = ops.get(op, 0) + 1
    top = sorted(ops.items(), key=lambda x: -x[1])[:4]
    if w > tw * 0.004 or sm > ts * 0.01:
        print(f"{s:5d} warp%={100*w/tw:5.1f} thr/warp={t/max(w,1):5.1f} samples%={100*sm/ts:5.1f} {top}")
