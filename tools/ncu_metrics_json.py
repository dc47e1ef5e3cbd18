"""Summary of an `ncu --set full` report as JSON (what profiles/ keeps): tools/ncu_metrics_json.py report.ncu-rep out.json
[--workload TEXT] [--build TEXT] [--alg-bytes N].  One entry per profiled kernel launch; dram_bytes_per_launch sums them."""
import argparse, csv, io, json, subprocess
ap = argparse.ArgumentParser()
ap.add_argument("report"); ap.add_argument("out")
ap.add_argument("--workload", default=""); ap.add_argument("--build", default=""); ap.add_argument("--alg-bytes", type=float, default=0)
a = ap.parse_args()
txt = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed_op_shared_atom.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
kernels, dram = [], 0.0
for r in rows[2:]:
    d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
    k = {"kernel": d.get("Kernel Name", "")}
    for m in KEEP:
        if m in d: k[m] = {"value": d[m], "unit": u.get(m, "")}
    for m in hdr:
        if "issue_stalled" in m and m.endswith("_per_issue_active.ratio"):
            try:
                if float(d[m]) > 0.1: k.setdefault("stall_cycles_per_issue", {})[m.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")] = float(d[m])
            except Exception: pass
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        try: dram += float(d[m]) * scale.get(u[m], 1)
        except Exception: pass
    kernels.append(k)
out = {"workload": a.workload, "build": a.build, "kernels": kernels, "dram_bytes_per_launch": dram}
if a.alg_bytes:
    out["algorithmic_bytes_per_launch"] = a.alg_bytes; out["traffic_over_algorithmic"] = dram / a.alg_bytes
json.dump(out, open(a.out, "w"), indent=1)
print(json.dumps({k: v for k, v in out.items() if k != "kernels"}))
