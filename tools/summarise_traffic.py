"""Per-kernel DRAM traffic / time from `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv`.

    python tools/summarise_traffic.py launches.csv out.json
"""
import csv
import json
import re
import sys
from collections import defaultdict


def main(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) >= 15 and r[0].isdigit()]
    per = defaultdict(dict)
    for r in rows:
        name = re.sub(r"\(.*", "", r[4]).replace("void ", "").replace("<unnamed>::", "")
        per[int(r[0])]["name"] = re.sub(r"^(ldm_gemm::)?(\w+)<.*", r"\2", name)  # template arguments folded together
        val = float(r[14].replace(",", ""))
        unit = r[13]
        if unit in ("Mbyte", "MB"):
            val *= 1e6
        elif unit in ("Kbyte", "KB"):
            val *= 1e3
        elif unit in ("Gbyte", "GB"):
            val *= 1e9
        elif unit in ("us", "usecond"):
            val *= 1e3
        elif unit in ("ms", "msecond"):
            val *= 1e6
        per[int(r[0])][r[12]] = val
    agg = defaultdict(lambda: {"launches": 0, "dram_bytes": 0.0, "ns": 0.0})
    for d in per.values():
        a = agg[d["name"]]
        a["launches"] += 1
        a["dram_bytes"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        a["ns"] += d.get("gpu__time_duration.sum", 0.0)
    res = {k: {"launches": v["launches"], "dram_bytes_per_launch": v["dram_bytes"] / v["launches"],
               "avg_us": v["ns"] / v["launches"] / 1e3, "total_ms": v["ns"] / 1e6} for k, v in agg.items()}
    json.dump(res, open(out, "w"), indent=1)
    for k, v in sorted(res.items(), key=lambda kv: -kv[1]["total_ms"]):
        print(f"{k[:48]:48s} n={v['launches']:4d} total={v['total_ms']:8.3f} ms  dram/launch={v['dram_bytes_per_launch'] / 1e6:8.2f} MB")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
