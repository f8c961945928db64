"""Code evidence for profiles/: the ptxas register / spill / stack table and a SASS mnemonic histogram of the hot kernels
of the library build() produces.   python tools/code_evidence.py [out_prefix]   (needs cuobjdump, c++filt; no GPU)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "fetal_t2mapping_b200", "csrc")
LIB = os.path.join(CSRC, "libt2fit.so")
HOT = ["fit_kernel<0, 5, 0, true>", "fit_kernel<0, 5, 0, false>", "fit_kernel<0, 6, 2, true>", "floor_queue_kernel<12, 0>",
       "floor_queue_kernel<16, 0>", "lbfgsb_dense_kernel<0>", "lbfgsb_dense_kernel<1>", "lbfgsb_dense_kernel<2>", "lbfgsb_kernel<0>", "lbfgsb_kernel<1>", "lbfgsb_kernel<2>", "lbfgsb_coop_kernel<1, 8>",
       "lbfgsb_coop_kernel<1, 32>", "zero_fill_kernel<true>", "zero_fill_kernel<false>", "mask_count_kernel", "mask_write_kernel",
       "mask_union_kernel", "roi_stats_kernel", "residual_kernel"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def short(d):
    d = re.sub(r"\(anonymous namespace\)::", "", d)
    d = re.sub(r"^void ", "", d)
    return re.sub(r"\(.*$", "", d)


def ptxas_table(out):
    rep = open(os.path.join(CSRC, "ptxas_report.txt")).read()
    ents = re.findall(r"Compiling entry function '([^']+)' for 'sm_100a'\n(?:.*\n)*?ptxas info\s+: Used (\d+) registers(.*)\n", rep)
    frames = dict(re.findall(r"Function properties for ([^\n]+)\n\s+(\d+ bytes stack frame, \d+ bytes spill stores, \d+ bytes spill loads)", rep))
    dm = demangle([e[0] for e in ents])
    rows = []
    for name, regs, rest in ents:
        s = short(dm[name])
        if any(s == h or s.startswith(h) for h in HOT):
            rows.append((s, int(regs), frames.get(name, ""), rest.strip(", ")))
    with open(out, "w") as f:
        f.write("ptxas -v (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo), hot kernels of libt2fit.so\n")
        f.write(f"{'kernel':44s} {'regs':>5s}  stack / spills; other\n")
        for s, r, fr, rest in sorted(rows):
            f.write(f"{s:44s} {r:5d}  {fr}; {rest}\n")
    return len(rows)


def sass_histogram(out):
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    blocks = re.split(r"\n\s+Function : ", sass)[1:]
    names = [b.split("\n", 1)[0].strip() for b in blocks]
    dm = demangle(names)
    with open(out, "w") as f:
        f.write("SASS mnemonic histogram (cuobjdump -sass of the built libt2fit.so; static instruction counts, callees of a kernel included)\n")
        for b, n in zip(blocks, names):
            s = short(dm[n])
            if not any(s == h for h in HOT[:13]):
                continue
            ops = collections.Counter()
            for m in re.finditer(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", b, re.M):
                ops[m.group(1).split(".")[0]] += 1
            tot = sum(ops.values())
            f.write(f"\n== {s}: {tot} instructions ({tot * 16 / 1024:.0f} KB)\n")
            f.write("   " + ", ".join(f"{k} {v}" for k, v in ops.most_common(28)) + "\n")
            tensor = [k for k in ops if k.startswith(("HMMA", "UTC", "TCGEN", "UTMA", "WGMMA"))]
            f.write(f"   tensor-core / TMA mnemonics: {tensor or 'none (scalar FP32 / FP64 + MUFU path, as north_star says)'}\n")


if __name__ == "__main__":
    pre = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02")
    n = ptxas_table(pre + "_ptxas_table.txt")
    sass_histogram(pre + "_sass_histogram.txt")
    print("kernels in the table:", n)
