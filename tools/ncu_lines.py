"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line."""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, hdr, out = None, None, []
for r in rows:
    if len(r) == 2 and r[0] in ("File Path", "File Name"):
        cur = r[1].split('/')[-1]
        continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0] != "":
        d = {}
        for k, v in zip(hdr, r):
            d.setdefault(k, v)
        try:
            out.append((int(d["Instructions Executed"]), int(d["# Samples"]), cur, r[0], r[1][:100],
                        d["Avg. Threads Executed"]))
        except Exception:
            pass
tot = sum(o[0] for o in out) or 1
tots = sum(o[1] for o in out) or 1
print("total warp-inst", tot, "samples", tots)
for o in sorted(out, reverse=True)[:top]:
    print("%5.1f%% inst %5.1f%% smp thr=%4s %s:%s  %s" % (100 * o[0] / tot, 100 * o[1] / tots, o[5], o[2], o[3], o[4]))
