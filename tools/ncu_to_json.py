"""One `ncu --set full` capture -> the JSON summary bench.py reads for its roofline keys (profiles/r02_<kernel>.json).

  python tools/ncu_to_json.py gpurun_out/r02_prof_extend.ncu-rep k_extend 8388608 "cmd line of the capture" > profiles/r02_k_extend.json

rays_per_launch: the jobs the captured launch processed (the capture commands fix it: pool_paths = 2^23 and a launch index at which
the stream is full). Everything else is read from the report's raw page."""
import csv
import io
import json
import subprocess
import sys

rep, label, rays = sys.argv[1], sys.argv[2], float(sys.argv[3])
cmd = sys.argv[4] if len(sys.argv) > 4 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, val = rows[0], rows[1], rows[2]
ix = {h: i for i, h in enumerate(hdr)}


def get(name, scale_to=None):
    if name not in ix:
        return None
    v, u = val[ix[name]].replace(",", ""), units[ix[name]]
    try:
        x = float(v)
    except ValueError:
        return None
    mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "second": 1.0, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9}
    return x * mult[u] if u in mult else x


dur = get("gpu__time_duration.sum")
winst = get("smsp__inst_executed.sum")
tinst = get("sass__thread_inst_executed_true_per_opcode") or get("smsp__thread_inst_executed.sum")
dr, dw = get("dram__bytes_read.sum") or 0.0, get("dram__bytes_write.sum") or 0.0
out = {
    "kernel": label, "kernel_name": val[ix["Kernel Name"]] if "Kernel Name" in ix else label, "capture": cmd, "report": rep.split("/")[-1],
    "rays_per_launch": rays, "duration_ms": dur * 1e3, "grays_per_s_under_ncu": rays / dur / 1e9,
    "registers_per_thread": get("launch__registers_per_thread"), "grid": get("launch__grid_size"), "block": get("launch__block_size"),
    "dynamic_smem_per_block": get("launch__shared_mem_per_block_dynamic"),
    "warps_active_pct": get("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "issue_active": (get("smsp__issue_active.avg.pct_of_peak_sustained_active") or 0) / 100.0,
    "warp_inst": winst, "warp_inst_per_ray": winst / rays, "threads_per_inst": (tinst / winst) if (tinst and winst) else None,
    "pipe_pct": {k: get(f"sm__inst_executed_pipe_{k}.avg.pct_of_peak_sustained_active") for k in ("fp64", "fma", "alu", "lsu", "xu")},
    "l1tex_hit_rate_pct": get("l1tex__t_sector_hit_rate.pct"), "lts_hit_rate_pct": get("lts__t_sector_hit_rate.pct"),
    "l1tex_frac": (get("l1tex__throughput.avg.pct_of_peak_sustained_elapsed") or 0) / 100.0,
    "lts_frac": (get("lts__throughput.avg.pct_of_peak_sustained_elapsed") or 0) / 100.0,
    "dram_frac_ncu": (get("dram__throughput.avg.pct_of_peak_sustained_elapsed") or 0) / 100.0,
    "dram_bytes_read": dr, "dram_bytes_write": dw, "dram_bytes_per_ray": (dr + dw) / rays, "dram_gbs": (dr + dw) / dur / 1e9,
    "stall_cycles_per_issue": {k: get(f"smsp__average_warps_issue_stalled_{k}_per_issue_active.ratio") for k in
                               ("long_scoreboard", "wait", "no_instruction", "short_scoreboard", "not_selected", "branch_resolving", "math_pipe_throttle",
                                "mio_throttle", "lg_throttle", "dispatch_stall", "barrier")},
}
print(json.dumps(out, indent=1))
