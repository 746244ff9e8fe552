"""Per-kernel DRAM traffic of the path kernels from an `ncu --set full` raw CSV (`ncu -i X.ncu-rep --page raw --csv`):
writes profiles/<tag>_ncu_traffic.json, which bench.py reads for `roofline.traffic`.

    python tools/ncu_traffic.py profiles/r02_ncu_full_raw.csv r02 131072 1
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNELS = {"conv1d": ("conv1d_fwd_kernel",), "ssd": ("ssd_fused_kernel", "ssd_dt_cumsum"), "gated_rmsnorm": ("gated_rmsnorm_kernel",)}


def main(path, tag, seqlen, n_gpus):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def bytes_of(r, col):
        v, u = float(r[ix[col]].replace(",", "")), units[ix[col]].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    per = {}
    for r in data:
        name = r[ix["Kernel Name"]]
        t = bytes_of(r, "dram__bytes_read.sum") + bytes_of(r, "dram__bytes_write.sum")
        per.setdefault(name, []).append(t)
    out = {}
    detail = {}
    for key, pats in KERNELS.items():
        tot = 0.0
        for name, vals in per.items():
            if any(p in name for p in pats):
                tot += sum(vals) / len(vals)
                detail[name[:80]] = {"launches_captured": len(vals), "dram_bytes_per_launch": sum(vals) / len(vals)}
        out[key] = tot
    blob = {"source": os.path.relpath(path, ROOT), "seqlen": int(seqlen), "n_gpus": int(n_gpus),
            "traffic_bytes_per_launch": out, "kernels": detail,
            "note": "dram__bytes_read.sum + dram__bytes_write.sum per launch; ssd = fused scan + dt/cumsum pre-kernel"}
    dst = os.path.join(ROOT, "profiles", f"{tag}_ncu_traffic.json")
    json.dump(blob, open(dst, "w"), indent=1)
    print(dst, out)


if __name__ == "__main__":
    main(*sys.argv[1:5])
