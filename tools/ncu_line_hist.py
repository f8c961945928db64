"""Executed warp instructions / active lanes / stall samples of one kernel by SOURCE LINE of a chosen file, from
`ncu --page source --csv` (SASS rows, addresses) joined with `nvdisasm -gi` of the same build (line + inline chain per
instruction; the library built here is the one that ran on the box).
    python tools/ncu_line_hist.py <source.csv> <kernel-symbol-substring> <file-substring> [first_line last_line]
An instruction is attributed to the OUTERMOST frame of its inline chain that lies in <file> (within [first_line, last_line]
when given), i.e. to the statement of that function it was inlined into; out-of-line callees (`$` sub-functions) are listed
by name."""
import collections
import csv
import os
import re
import subprocess
import sys

INNER = bool(os.environ.get("LINE_HIST_INNER"))      # attribute to the INNERMOST frame in <file>, grouped by function
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.environ.get("T2FIT_LIB", os.path.join(ROOT, "fetal_t2mapping_b200", "csrc", "libt2fit.so"))


def disasm(sym):
    tmp = "/tmp/ncu_line_hist"
    os.makedirs(tmp, exist_ok=True)
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, check=True, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
    out, inside = [], False
    for ln in txt:
        if ln.startswith("//--------------------- .text"):
            inside = sym in ln
        if inside:
            out.append(ln)
    return out


def parse(lines, fsub, lo, hi):
    """offset -> (label, innermost 'file:line')"""
    table, chain, func = {}, [], "kernel"
    for ln in lines:
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
        if m:
            if not chain or chain[-1][2] != (m.group(1), int(m.group(2))):
                chain = []
            chain.append(((m.group(1), int(m.group(2))), None, (m.group(3), int(m.group(4))) if m.group(3) else None))
            continue
        m = re.match(r"^(\$?[\w$]+):\s*$", ln)
        if m and "$" in m.group(1):
            func = m.group(1).split("$")[-1]
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            off = int(m.group(1), 16)
            frames = [c[0] for c in chain]
            if chain and chain[-1][2]:
                frames.append(chain[-1][2])
            label = None
            order = frames if INNER else list(reversed(frames))     # innermost / outermost frame in <file> first
            for f, l in order:
                if fsub in f and lo <= l <= hi:
                    label = l
                    break
            inner = f"{os.path.basename(frames[0][0])}:{frames[0][1]}" if frames else "?"
            table[off] = (func, label, inner)
            chain = chain if False else chain      # the chain stays in force until the next //## block
    return table


def main():
    path, sym, fsub = sys.argv[1:4]
    lo, hi = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (0, 1 << 30)
    table = parse(disasm(sym), fsub, lo, hi)
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ia, iex, ithr, ismp = (hdr.index(n) for n in ("Address", "Instructions Executed", "Thread Instructions Executed", "# Samples"))
    ist = [hdr.index(n) for n in ("stall_long_sb", "stall_no_inst", "stall_wait", "stall_short_sb", "stall_branch_resolving")]
    base = None
    agg = collections.defaultdict(lambda: [0] * 9)
    tot = [0, 0, 0]
    for r in rows[2:]:
        if len(r) <= ismp or not r[ia].startswith("0x"):
            continue
        a = int(r[ia], 16)
        base = a if base is None else base
        func, label, inner = table.get(a - base, ("?", None, "?"))
        key = f"{func}" if func != "kernel" else (f"line {label}" if label is not None else f"other ({inner.split(':')[0]})")
        g = agg[key]
        g[0] += int(r[iex]); g[1] += int(r[ithr]); g[2] += int(r[ismp]); g[3] += 1
        for q, col in enumerate(ist):
            g[4 + q] += int(r[col] or 0)
        tot[0] += int(r[iex]); tot[1] += int(r[ithr]); tot[2] += int(r[ismp])
    src = {}
    try:
        cand = [os.path.join(dp, f) for dp, _, fs in os.walk(ROOT) for f in fs if fsub in f and not dp.startswith(os.path.join(ROOT, "gpurun_out"))]
        src = dict(enumerate(open(cand[0]).read().splitlines(), 1)) if cand else {}
    except OSError:
        pass
    if INNER and src:                                # group lines by the function that contains them
        starts = [(n, re.search(r"(\w+)\s*\(", t).group(1)) for n, t in src.items()
                  if re.match(r"\s*(template\s*<[^>]*>\s*)?(T2_HD|T2_NI|T2_LP|__device__|static|inline)\b.*\w+\s*\(.*\)\s*(const)?\s*\{", t)]
        def fn_of(l):
            name = "?"
            for n, nm in starts:
                if n <= l:
                    name = nm
            return name
        agg2 = collections.defaultdict(lambda: [0] * 9)
        for key, g in agg.items():
            m = re.match(r"line (\d+)", key)
            k2 = f"fn {fn_of(int(m.group(1)))}" if m else key
            for i in range(9):
                agg2[k2][i] += g[i]
        agg = agg2
    print(f"total: {tot[0]} warp instructions, {tot[1] / max(tot[0], 1):.1f} lanes, {tot[2]} samples")
    print(f"{'where':28s} {'warp instr':>12s} {'share':>6s} {'lanes':>6s} {'samples':>8s} {'SASS':>5s} {'long_sb':>8s} {'no_inst':>8s} {'wait':>6s} {'short':>6s} {'branch':>6s}  source   (stall columns: % of ALL samples)")
    for key, g in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
        text = ""
        m = re.match(r"line (\d+)", key)
        if m:
            text = src.get(int(m.group(1)), "").strip()[:90]
        print(f"{key[-28:]:28s} {g[0]:12d} {100.0 * g[0] / tot[0]:5.1f}% {g[1] / max(g[0], 1):6.1f} {100.0 * g[2] / max(tot[2], 1):7.1f}% {g[3]:5d} {100.0 * g[4] / max(tot[2], 1):7.1f}% {100.0 * g[5] / max(tot[2], 1):7.1f}% {100.0 * g[6] / max(tot[2], 1):5.1f}% {100.0 * g[7] / max(tot[2], 1):5.1f}% {100.0 * g[8] / max(tot[2], 1):5.1f}%  {text}")


if __name__ == "__main__":
    main()
