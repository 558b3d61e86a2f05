"""Samples per CUDA source line from `ncu --page source --print-source cuda,sass --csv` (first launch of the kernel).

    ncu -i X.ncu-rep --page source --print-source cuda,sass --csv --kernel-name regex:NAME > /tmp/src.csv
    python tools/ncu_source_summary.py /tmp/src.csv [topN] [kernel-substring]
"""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    want = sys.argv[3] if len(sys.argv) > 3 else None
    rows = list(csv.reader(open(path)))
    secs = [i for i, r in enumerate(rows) if r and r[0] == "File Path"]
    agg, srcs, stall = defaultdict(int), {}, defaultdict(lambda: defaultdict(int))
    done = set()
    for si, s in enumerate(secs):
        fp, fn = rows[s][1], rows[s + 1][1]
        if want and want not in fn:
            continue
        if (fp, fn) in done:
            continue  # later launches of the same kernel
        done.add((fp, fn))
        hdr = rows[s + 2]
        end = secs[si + 1] if si + 1 < len(secs) else len(rows)
        col = {}
        for i, h in enumerate(hdr):
            col.setdefault(h, i)
        scols = [h for h in col if h.startswith("stall_") and "Not Issued" not in h]
        for r in rows[s + 3:end]:
            if len(r) < len(hdr) or r[0] == "":
                continue
            try:
                ns = int(r[col["# Samples"]])
            except ValueError:
                continue
            k = (fp.split("/")[-1], int(r[0]))
            srcs[k] = r[1].strip()
            agg[k] += ns
            for h in scols:
                try:
                    stall[k][h[6:]] += int(r[col[h]])
                except ValueError:
                    pass
    tot = sum(agg.values())
    print(f"total samples {tot}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:topn]:
        st = sorted(stall[k].items(), key=lambda kv: -kv[1])[:3]
        print(f"{v:7d} {100 * v / max(1, tot):5.1f}%  {k[0]}:{k[1]:<4d} {srcs[k][:95]:95s} {st}")


if __name__ == "__main__":
    main()
