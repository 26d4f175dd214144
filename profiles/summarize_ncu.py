"""Summarise an ncu metrics CSV of ONE all-pairs step (tests/gpu_perf_probe.py once 1024 128:2) into
profiles/r02_step_kernels.summary.txt and profiles/r02_traffic.json (the figure bench.py quotes as roofline.traffic,
keyed by a hash of the kernel sources so that a stale capture is never quoted).

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_active.avg \\
        --clock-control none --csv --log-file gpurun_out/r02_step_metrics.csv python tests/gpu_perf_probe.py once 1024 128:2
    python profiles/summarize_ncu.py gpurun_out/r02_step_metrics.csv 1024
"""
import collections
import csv
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def csrc_sha():
    h = hashlib.sha256()
    d = os.path.join(ROOT, "clip_embeds_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh")):
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def main():
    path, images = sys.argv[1], int(sys.argv[2])
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, mi, vi, ui, idi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
    launches = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ki]
        if not any(t in name for t in ("clipk::", "eng2::", "eng::", "fz::", "mega::")):
            continue                                   # the probe's own input generation / torch copies
        d = launches.setdefault(r[idi], {"name": name})
        v = float(r[vi].replace(",", ""))
        unit = r[ui]
        if "bytes" in r[mi]:
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        if "time_duration" in r[mi]:
            v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3}.get(unit, 1e-3)
        d[r[mi]] = v
    agg = collections.OrderedDict()
    for d in launches.values():
        short = d["name"].split("(")[0]
        short = short.replace("void ", "")
        a = agg.setdefault(short, {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0, "tensor": []})
        a["n"] += 1
        a["us"] += d.get("gpu__time_duration.sum", 0.0)
        a["rd"] += d.get("dram__bytes_read.sum", 0.0)
        a["wr"] += d.get("dram__bytes_write.sum", 0.0)
        for k, v in d.items():
            if "tensor" in k:
                a["tensor"].append(v)
    tot_us = sum(a["us"] for a in agg.values())
    tot_b = sum(a["rd"] + a["wr"] for a in agg.values())
    lines = [f"# one all-pairs step (B = {images}, P = 576, D = 768, bf16), ncu --clock-control none, serialised launches (cold caches):",
             f"# compare SHARES, not absolute times.  source csv: {os.path.basename(path)}   kernel sources sha16 {csrc_sha()}",
             f"{'kernel':<78} {'n':>3} {'us':>9} {'share':>6} {'dram rd MB':>11} {'dram wr MB':>11} {'tensor pipe %':>13}"]
    for k, a in agg.items():
        t = (sum(a["tensor"]) / len(a["tensor"])) if a["tensor"] else float("nan")
        lines.append(f"{k[:78]:<78} {a['n']:>3} {a['us']:>9.1f} {100 * a['us'] / tot_us:>5.1f}% {a['rd'] / 1e6:>11.1f} {a['wr'] / 1e6:>11.1f} {t:>13.1f}")
    lines.append(f"{'TOTAL':<78} {sum(a['n'] for a in agg.values()):>3} {tot_us:>9.1f} 100.0% {sum(a['rd'] for a in agg.values()) / 1e6:>11.1f} "
                 f"{sum(a['wr'] for a in agg.values()) / 1e6:>11.1f}")
    lines.append(f"# dram bytes per step: {tot_b / 1e9:.3f} GB   (algorithmic minimum 3 B P D 2 = {3 * images * 576 * 768 * 2 / 1e9:.3f} GB)")
    out = os.path.join(ROOT, "profiles", "r02_step_kernels.summary.txt")
    open(out, "w").write("\n".join(lines) + "\n")
    json.dump({"csrc_sha16": csrc_sha(), "images": images, "dram_bytes_per_step": tot_b,
               "source": "profiles/r02_step_kernels.summary.txt (ncu dram__bytes_read.sum + dram__bytes_write.sum over every kernel of one step)"},
              open(os.path.join(ROOT, "profiles", "r02_traffic.json"), "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
