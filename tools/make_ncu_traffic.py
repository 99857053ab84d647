#!/usr/bin/env python
"""profiles/ncu_traffic.json from an `ncu --set full ... --page raw --csv` dump of `bench.py --steps 1 --warmup 1`:
DRAM bytes (read + write) of the LONGEST launch of the GEMM class and of the conv class, keyed by the sha256 of the kernel
sources so that bench.py reports `roofline.traffic` only for the build the capture belongs to.

    python tools/make_ncu_traffic.py gpurun_out/<tag>_full_raw.csv [workload] [gemm_backend]"""
import csv
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aimnetcentral_b200 import build  # noqa: E402

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TIME = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3}


def main():
    path = sys.argv[1]
    workload = sys.argv[2] if len(sys.argv) > 2 else "cfg2"
    backend = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    rows = list(csv.reader(open(path, newline="")))
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {k: hdr.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum")}
    best = {}
    for r in body:
        name = r[col["Kernel Name"]]
        cls = "gemm" if "gemm_tc16" in name else ("conv" if ("conv_" in name or "fwd_kernel" in name or "bwd_kernel" in name) and "prep" not in name else None)
        if cls is None:
            continue
        val = lambda k: float(r[col[k]].replace(",", ""))   # noqa: E731
        us = val("gpu__time_duration.sum") * TIME[units[col["gpu__time_duration.sum"]]]
        rd = val("dram__bytes_read.sum") * UNIT[units[col["dram__bytes_read.sum"]]]
        wr = val("dram__bytes_write.sum") * UNIT[units[col["dram__bytes_write.sum"]]]
        if cls not in best or us > best[cls]["us_under_ncu"]:
            best[cls] = {"kernel": name.split("(")[0], "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes": rd + wr,
                         "us_under_ncu": us}
    out = {"build_sha256": build._digest(), "workload": workload, "gemm_backend": backend, "source": os.path.basename(path),
           "note": "longest launch of each class in one ncu --set full capture of bench.py --steps 1 --warmup 1", **best}
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
