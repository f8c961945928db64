"""Executed warp instructions and stall samples per device function of one kernel:
   python tools/ncu_func_hist.py <source.csv from `ncu --page source --csv`> <lib.so> <kernel-name-substring>"""
import csv
import re
import subprocess
import sys


def main(src_csv, lib, kname):
    rows = list(csv.reader(open(src_csv)))
    hdr = rows[1]
    ia, iex, ithr, ismp = (hdr.index(n) for n in ("Address", "Instructions Executed", "Thread Instructions Executed", "# Samples"))
    ins = [(int(r[ia], 16), int(r[iex]), int(r[ithr]), int(r[ismp])) for r in rows[2:] if len(r) > ismp and r[ia].startswith("0x")]
    base = ins[0][0]
    elf = subprocess.run(["cuobjdump", "-elf", lib], capture_output=True, text=True).stdout
    funcs = []
    for line in elf.splitlines():
        f = line.split()
        if len(f) >= 7 and f[0].startswith("0x") and kname in f[-1]:
            try:
                off, size = int(f[1], 16), int(f[2], 16)
            except ValueError:
                continue
            name = f[-1].split("$")[-1] if "$" in f[-1] else "(kernel body)"
            funcs.append((off, size, name))
    funcs.sort()
    tot = sum(i[1] for i in ins)
    out = []
    for off, size, name in funcs:
        sel = [i for i in ins if off <= i[0] - base < off + size] if name != "(kernel body)" else []
        out.append((sum(i[1] for i in sel), sum(i[2] for i in sel), sum(i[3] for i in sel), name))
    first = min(o for o, _, n in funcs if n != "(kernel body)")
    sel = [i for i in ins if i[0] - base < first]
    out.append((sum(i[1] for i in sel), sum(i[2] for i in sel), sum(i[3] for i in sel), "(kernel body)"))
    print(f"total warp instructions {tot}")
    for n, t, s, name in sorted(out, reverse=True):
        if n:
            print(f"  {100.0 * n / tot:5.1f} %  lanes {t / n:5.1f}  samples {s:7d}  {name[:90]}")


if __name__ == "__main__":
    main(*sys.argv[1:4])
