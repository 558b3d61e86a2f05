"""Per-SASS-instruction warp-stall samples from `ncu -i X.ncu-rep --page source --print-source sass --csv > f.csv`.

    python tools/ncu_sass_hot.py f.csv [min_samples] > listing.txt
Prints address offset, samples, executions, instruction and the three largest stall reasons.
"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    thr = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    hdr = rows[1]
    stall = [(h[6:], i) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    base = None
    tot = 0
    for r in rows[2:]:
        if not r or not r[0].startswith("0x"):
            continue
        a = int(r[0], 16)
        base = a if base is None else base
        n = int(r[4] or 0)
        tot += n
        if n < thr:
            continue
        st = sorted(((int(r[i] or 0), h) for h, i in stall if i < len(r)), reverse=True)[:3]
        print(f"{a - base:05x} {n:6d} {r[5]:>9} {r[1].strip()[:64]:64s} {[(h, c) for c, h in st if c > 0]}")
    print("total samples", tot)


if __name__ == "__main__":
    main()
