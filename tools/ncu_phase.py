"""Per-phase shares (instructions, samples) of the mask kernel from an ncu report: tools/ncu_phase.py report.ncu-rep
Phases are found by the marker comments of the working copy of ta_scan_mask.cuh (the report must come from that source)."""
import csv, subprocess, sys, io, re
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur = None; hdr = None; out = []; src = {}
for r in rows:
    if len(r) == 2 and r[0] in ("File Path", "File Name"): cur = r[1].split('/')[-1]; continue
    if len(r) > 5 and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0] != "":
        d = {}
        for k, v in zip(hdr, r): d.setdefault(k, v)
        try:
            out.append((cur, int(r[0]), int(d["# Samples"]), int(d["Instructions Executed"]), d))
            if cur == 'ta_scan_mask.cuh': src[int(r[0])] = r[1]
        except Exception: pass
marks = [("helpers/init", r"^// The streaming pass"), ("stage", r"---- stage the tile"), ("P1", r"---- P1:"), ("barrier1", r"__syncthreads_and\(one_label\)"),
         ("between", r"const bool overflow = "), ("P2 setup", r"---- P2:"), ("P2 mom", r"if \(do_mom\) \{"), ("P2 pairs", r"if \(do_pairs\) \{"),
         ("G", r"---- G:"), ("barrier2", r"MK_TICK\(3\)"), ("F", r"---- F:"), ("clear", r"back to all-zero masks")]
starts = []
import os
full = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tissue_analysis_b200", "csrc", "ta_scan_mask.cuh")).read().split("\n")
for name, pat in marks:
    for ln, text in enumerate(full, 1):
        if re.search(pat, text) and (not starts or ln > starts[-1][1]): starts.append((name, ln)); break
ts = sum(o[2] for o in out) or 1; ti = sum(o[3] for o in out) or 1
acc = {}; files = {}
for f, l, s, i, d in out:
    if f == 'ta_scan_mask.cuh':
        name = "?"
        for n, a in starts:
            if l >= a: name = n
        acc.setdefault(name, [0, 0]); acc[name][0] += s; acc[name][1] += i
    else:
        files.setdefault(f, [0, 0]); files[f][0] += s; files[f][1] += i
print("total warp-inst %d, samples %d" % (ti, ts))
for n, a in starts:
    v = acc.get(n, [0, 0]); print("%-14s from line %4d: inst %5.1f%%  samples %5.1f%%" % (n, a, 100 * v[1] / ti, 100 * v[0] / ts))
for f, v in files.items(): print("%-36s inst %5.1f%%  samples %5.1f%%" % (f, 100 * v[1] / ti, 100 * v[0] / ts))
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
if lo:
    hi = int(sys.argv[3])
    for f, l, s, i, d in sorted(out, key=lambda o: o[1]):
        if f == 'ta_scan_mask.cuh' and lo <= l <= hi and i: print("%4d %5.2f%% inst %5.2f%% smp | %s" % (l, 100 * i / ti, 100 * s / ts, src[l][:110]))
