"""Opcode histogram (weighted by executed warp instructions) and hottest stall sites from `ncu --page source --csv`."""
import csv
import collections
import re
import sys


def main(path, top=25):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ia, isrc, iex, ithr, ismp = (hdr.index(n) for n in ("Address", "Source", "Instructions Executed",
                                                          "Thread Instructions Executed", "# Samples"))
    ops = collections.Counter()
    thr = collections.Counter()
    total = 0
    lines = []
    for r in rows[2:]:
        if len(r) <= ismp or not r[ia].startswith("0x"):
            continue
        src = r[isrc].strip()
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", src)
        op = m.group(2) if m else src.split()[0]
        n = int(r[iex]); t = int(r[ithr])
        ops[op] += n; thr[op] += t; total += n
        lines.append((int(r[ismp]), n, src))
    print(f"total executed warp instructions: {total}")
    for op, n in ops.most_common(top):
        print(f"  {op:10s} {n:10d}  {100.0 * n / total:5.1f} %   avg lanes {thr[op] / max(n, 1):5.1f}")
    print("hottest by stall samples:")
    for smp, n, src in sorted(lines, reverse=True)[:15]:
        print(f"  samples {smp:6d}  executed {n:9d}  {src[:80]}")


if __name__ == "__main__":
    main(sys.argv[1])
