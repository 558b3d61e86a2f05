"""Summarise `ncu --page source --print-source cuda,sass --csv` output: samples per CUDA source line and per stall reason.

    ncu -i X.ncu-rep --page source --print-source cuda,sass --csv --kernel-name regex:NAME | python tools/ncu_source_summary.py [N]
"""
import csv
import sys


def main():
    topn = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    rows = list(csv.reader(sys.stdin))
    his = [i for i, r in enumerate(rows) if r and r[0] in ("Address", "#", "Line") or (r and "# Samples" in r)]
    if not his:
        print("no table found")
        return
    # only the first kernel instance
    blocks = []
    for k, hi in enumerate(his):
        end = his[k + 1] - 1 if k + 1 < len(his) else len(rows)
        blocks.append((rows[hi], rows[hi + 1:end]))
    for hdr, data in blocks[:2]:
        col = {h: i for i, h in enumerate(hdr)}
        if "# Samples" not in col:
            continue
        stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        tot = {h: 0 for h in stall_cols}
        recs = []
        for r in data:
            if len(r) < len(hdr):
                continue
            try:
                ns = int(r[col["# Samples"]])
            except ValueError:
                continue
            recs.append((ns, r))
            for h in stall_cols:
                try:
                    tot[h] += int(r[col[h]])
                except ValueError:
                    pass
        total = sum(ns for ns, _ in recs)
        print(f"== view with first column '{hdr[0]}': {len(recs)} rows, {total} samples")
        print("   stalls:", ", ".join(f"{h[6:]}={v}" for h, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]))
        for ns, r in sorted(recs, key=lambda t: -t[0])[:topn]:
            st = sorted(((h[6:], int(r[col[h]])) for h in stall_cols if r[col[h]] not in ("", "0")), key=lambda kv: -kv[1])[:3]
            src = r[col["Source"]].strip()[:100]
            print(f"{ns:8d} {100.0 * ns / max(1, total):5.1f}% ex={r[col['Instructions Executed']]:>10s} {r[0][-6:]:>6s} {src}  {st}")


if __name__ == "__main__":
    main()
