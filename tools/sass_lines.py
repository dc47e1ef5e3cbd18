"""Static SASS statistics per source line of one kernel (no GPU needed).

    cuobjdump -xelf all build/libtissue_b200_block.so          # -> ta_api.sm_100a.cubin
    nvdisasm -g -c ta_api.sm_100a.cubin > all.sass
    python tools/sass_lines.py all.sass scan_level_kernelItLb1 [--ranges FILE:LO-HI=name ...]

Prints the instruction count of the kernel, its local-memory (spill) instructions, the opcode mix and the instruction
count per source line (innermost inlined line as nvdisasm -g reports it), optionally summed over named line ranges.  A
static count is not a dynamic one: unrolled straight-line code counts once per copy, loops count once.
"""
import collections
import re
import sys


def main():
    path, pattern = sys.argv[1], sys.argv[2]
    ranges = []
    for a in sys.argv[3:]:
        if a.startswith("--"):
            continue
        spec, name = a.split("=")
        f, lohi = spec.split(":")
        lo, hi = lohi.split("-")
        ranges.append((f, int(lo), int(hi), name))
    infun = False
    cur = ("?", 0)
    per_line = collections.Counter()
    spill = collections.Counter()
    ops = collections.Counter()
    total = 0
    fun = None
    for line in open(path, errors="replace"):
        if line.startswith(".text."):
            infun = pattern in line
            if infun:
                fun = line.strip().rstrip(":")
            continue
        if not infun:
            continue
        m = re.match(r'\s*//## File "(.*)", line (\d+)', line)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        total += 1
        per_line[cur] += 1
        ops[op.split(".")[0]] += 1
        if op.startswith("STL") or op.startswith("LDL"):
            spill[cur] += 1
    print(fun)
    print("instructions: %d, local-memory instructions: %d" % (total, sum(spill.values())))
    print("opcode mix:", ", ".join("%s %d" % kv for kv in ops.most_common(24)))
    if ranges:
        for f, lo, hi, name in ranges:
            n = sum(c for (ff, l), c in per_line.items() if ff == f and lo <= l <= hi)
            s = sum(c for (ff, l), c in spill.items() if ff == f and lo <= l <= hi)
            print("  %-28s %s:%d-%d  %6d instructions, %4d local-memory" % (name, f, lo, hi, n, s))
    print("top lines:")
    for (f, l), c in per_line.most_common(40):
        print("  %-22s %5d  %6d  (local %d)" % (f, l, c, spill[(f, l)]))
    print("local-memory instructions by line:")
    for (f, l), c in spill.most_common(25):
        print("  %-22s %5d  %6d" % (f, l, c))


if __name__ == "__main__":
    try:
        main()
    except BrokenPipeError:
        pass
