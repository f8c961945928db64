"""profiles/traffic.json from an `ncu --set full ... --page raw --csv` export of one bench step:
DRAM bytes (read + write) per launch of the fit kernel and of the zero-fill kernel.
    python tools/make_traffic.py gpurun_out/step_raw.csv"""
import csv
import re
import json
import os
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ir, iw, ik, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name"), hdr.index("gpu__time_duration.sum")
    per = {}
    for r in rows[2:]:
        name = r[ik]
        # FILL = true (4th template argument): the launch also zero-fills the dense maps
        fused = re.search(r"fit_kernel<[^>]*,\s*(\(bool\))?(1|true)>", name) is not None
        key = "fused_fit_fill_kernel" if fused else "fit_kernel" if "fit_kernel" in name else "zero_fill_kernel" if "zero_fill" in name else "lbfgsb_kernel" if "lbfgsb" in name else None
        if key is None:
            continue
        b = float(r[ir]) * UNIT[units[ir]] + float(r[iw]) * UNIT[units[iw]]
        per.setdefault(key, []).append((b, float(r[it])))
    out = {"source": os.path.basename(path), "note": "ncu --set full --clock-control none, one launch each (cold cache, serialised)"}
    for k, v in per.items():
        out[f"{k}_dram_bytes_per_launch"] = sum(x[0] for x in v) / len(v)
        out[f"{k}_ncu_time_{units[it]}"] = sum(x[1] for x in v) / len(v)
    if "fused_fit_fill_kernel" in per:
        out["step_dram_bytes"] = out["fused_fit_fill_kernel_dram_bytes_per_launch"]
    elif "fit_kernel" in per and "zero_fill_kernel" in per:
        out["step_dram_bytes"] = out["fit_kernel_dram_bytes_per_launch"] + out["zero_fill_kernel_dram_bytes_per_launch"]
    json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1])
