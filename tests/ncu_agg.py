"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (diagnostic helper)."""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    tot = 0.0
    for row in csv.DictReader(lines):
        name = row["Kernel Name"]
        name = re.sub(r"void eng::gemm_kernel<\(?(?:int\))?(\d+), \(?(?:bool\))?(\d), \(?(?:bool\))?(\d), epi::(\w+(?:<[^>]*>)?)>.*",
                      r"gemm<BN=\1,Amn=\2,Bmn=\3,\4>", name)[:80]
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{t / 1e3:9.3f} ms {n:5d} launches avg {t / n:8.1f} us {100 * t / tot:5.1f}%  {k}")
    print(f"total {tot / 1e3:.3f} ms")


if __name__ == "__main__":
    main(sys.argv[1])
