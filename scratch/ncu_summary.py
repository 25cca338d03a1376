"""Summarise an `ncu --set full` report into profiles/: one CSV row per captured kernel launch and a JSON with the
per-kernel figures bench.py quotes (DRAM traffic per launch, tensor-pipe utilisation).
usage: python scratch/ncu_summary.py REPORT.ncu-rep OUT_PREFIX BATCH"""
import csv, io, json, subprocess, sys
rep, prefix, B = sys.argv[1], sys.argv[2], int(sys.argv[3])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr = rows[0]; units = rows[1]
KEEP = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__inst_executed_pipe_uniform.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]
cols = [k for k in KEEP if k in hdr]
out = io.StringIO(); w = csv.writer(out)
w.writerow(cols); w.writerow([units[hdr.index(k)] for k in cols])
summary = {}
def num(d, k):
    try: return float(d[k].replace(",", ""))
    except Exception: return None
def unit_scale(k):
    u = units[hdr.index(k)]
    return {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(u, 1.0)
for r in rows[2:]:
    d = dict(zip(hdr, r))
    w.writerow([d[k] for k in cols])
    name = d["Kernel Name"]
    if name in summary: continue
    rd = num(d, "dram__bytes_read.sum"); wr = num(d, "dram__bytes_write.sum")
    summary[name] = {
        "batch": B, "time_us_under_ncu": num(d, "gpu__time_duration.sum"),
        "dram_bytes_read": None if rd is None else rd * unit_scale("dram__bytes_read.sum"),
        "dram_bytes_write": None if wr is None else wr * unit_scale("dram__bytes_write.sum"),
        "tensor_pipe_pct_of_peak": num(d, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        "ipc": num(d, "sm__inst_executed.avg.per_cycle_active"), "warp_instructions": num(d, "smsp__inst_executed.sum"),
        "registers": num(d, "launch__registers_per_thread"), "warps_active_pct": num(d, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        "smem_wavefronts": num(d, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
        "smem_bank_conflict_wavefronts": num(d, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    }
open(prefix + ".csv", "w").write(out.getvalue())
json.dump(summary, open(prefix + ".json", "w"), indent=1)
for k, v in summary.items():
    print(k[:90], {a: b for a, b in v.items() if a != "batch"})
