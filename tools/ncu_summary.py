"""Key numbers of one kernel from an .ncu-rep (raw page): time, instructions, pipes, shared-memory wavefronts, stalls.
usage: python tools/ncu_summary.py rep.ncu-rep [units_per_launch]   (units = e.g. row-pair frames, for per-unit figures)"""
import csv, subprocess, sys, io
rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, v = rows[0], rows[2]
d = dict(zip(h, v))
def g(k):
    try: return float(d[k].replace(",", ""))
    except Exception: return None
keys = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for k in keys:
    if k in d: print(f"{k:75s} {d[k]:>16s} {rows[1][h.index(k)]}")
if units:
    print(f"per unit: instr {g('smsp__inst_executed.sum') / units:.0f}, shared-memory wavefronts {g('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum') / units:.0f}")
st = sorted(((float(v[i]), n.replace("smsp__average_warp_latency_issue_stalled_", "").replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
             for i, n in enumerate(h) if "issue_stalled" in n and n.endswith("_per_issue_active.ratio") and "not_issued" not in n), reverse=True)
print("stall cycles per issued instruction:", ", ".join(f"{n} {x:.2f}" for x, n in st[:10]))
