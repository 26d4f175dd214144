"""Per-kernel stall summary from `ncu -i X.ncu-rep --page source --csv` output (diagnostic helper)."""
import csv
import sys


def main(path, which=None, ntop=25):
    blocks, cur = [], None
    for row in csv.reader(open(path)):
        if row and row[0] == "Kernel Name":
            cur = {"name": row[1], "rows": [], "hdr": None}
            blocks.append(cur)
            continue
        if cur is None:
            continue
        if cur["hdr"] is None:
            cur["hdr"] = row
            continue
        cur["rows"].append(row)
    for bi, b in enumerate(blocks):
        if which is not None and bi != which:
            continue
        h = b["hdr"]
        si, src, ie = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
        stall_cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
        tot = sum(int(r[si]) for r in b["rows"])
        print(f"==== [{bi}] {b['name'][:90]}  total samples {tot}")
        for i in stall_cols:
            s = sum(int(r[i]) for r in b["rows"])
            if s > tot * 0.02:
                print(f"   {h[i]:28s} {s:6d} {100 * s / tot:5.1f}%")
        for r in sorted(b["rows"], key=lambda r: -int(r[si]))[:ntop]:
            reasons = sorted([(int(r[i]), h[i][6:]) for i in stall_cols], reverse=True)[:2]
            print(f"{int(r[si]):6d} {r[ie]:>8} {r[src].strip()[:72]:72s} {reasons}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else None, int(sys.argv[3]) if len(sys.argv) > 3 else 25)
