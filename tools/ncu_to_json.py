"""Write the profiles/*.json summary of one kernel from an .ncu-rep (raw page).
usage: python tools/ncu_to_json.py rep.ncu-rep out.json key=value ...   (extra key=value pairs are copied into the JSON)"""
import csv, io, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
extra = dict(kv.split("=", 1) for kv in sys.argv[3:])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
d = dict(zip(rows[0], rows[2])); u = dict(zip(rows[0], rows[1]))
def g(k, scale=1.0):
    try: return float(d[k].replace(",", "")) * scale
    except Exception: return None
def mb(k):
    v = g(k)
    if v is None: return None
    return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u.get(k, "byte"), 1.0)
t = g("gpu__time_duration.sum"); t = t * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u.get("gpu__time_duration.sum", "us"), 1.0) if t else None
res = {
    "kernel": d.get("Kernel Name"), "gpu_time_us_under_ncu": t,
    "dram_read_MB": mb("dram__bytes_read.sum"), "dram_write_MB": mb("dram__bytes_write.sum"),
    "dram_bytes_per_launch": int(1e6 * ((mb("dram__bytes_read.sum") or 0) + (mb("dram__bytes_write.sum") or 0))),
    "warp_instructions": g("smsp__inst_executed.sum"),
    "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "fma_pipe_active_pct": g("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
    "lsu_data_pipe_wavefronts_pct": g("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
    "tensor_pipe_active_pct": g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    "registers_per_thread": g("launch__registers_per_thread"), "grid": g("launch__grid_size"), "block": g("launch__block_size"),
    "dynamic_smem_KB": g("launch__shared_mem_per_block_dynamic"), "warps_active_pct": g("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "smem_wavefronts": g("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
    "smem_bank_conflict_wavefronts": g("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    "l1_hit_pct": g("l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": g("lts__t_sector_hit_rate.pct"),
    "stall_cycles_per_issue": {n.replace("smsp__average_warps_issue_stalled_", "").replace("smsp__average_warp_latency_issue_stalled_", "").replace("_per_issue_active.ratio", ""): round(float(v), 3)
                               for n, v in d.items() if "issue_stalled" in n and n.endswith("_per_issue_active.ratio") and "not_issued" not in n and v not in ("", "n/a")},
}
res.update(extra)
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1)[:600])
