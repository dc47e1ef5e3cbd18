"""Key numbers of one ncu report: tools/ncu_summary.py report.ncu-rep"""
import csv, io, subprocess, sys
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[0]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print(d.get("Kernel Name", "")[:60])
    for k in ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
              "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
              "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
              "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_atom.sum",
              "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_lsu.sum",
              "sm__inst_executed_pipe_uniform.sum", "sm__inst_executed_pipe_adu.sum", "sm__inst_executed_pipe_cbu.sum"]:
        if k in d: print("  %-70s %s" % (k, d[k]))
    for k in hdr:
        if "issue_stalled" in k and "per_issue_active" in k:
            try:
                if float(d[k]) > 0.15: print("   stall %-30s %s" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), d[k]))
            except Exception: pass
