"""Aggregates an ncu launch list (--csv --log-file, metrics gpu__time_duration.sum [+ dram__bytes_*]) by kernel:
launches, total / average duration, DRAM bytes read + written.  Prints a TSV and the per-step totals."""
import collections
import csv
import re
import sys


def load(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hdr = None
    per = collections.OrderedDict()
    for r in rows:
        if r and r[0] == "ID":
            hdr = r
            continue
        if not hdr or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        key = d["ID"]
        k = per.setdefault(key, {"name": re.sub(r"\(.*", "", d["Kernel Name"]).replace("void ", "").replace("inb::", "")})
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit, m = d["Metric Unit"], d["Metric Name"]
        if m == "gpu__time_duration.sum":
            k["us"] = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
        elif m.startswith("dram__bytes"):
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
            k["dram"] = k.get("dram", 0.0) + v * scale
    return list(per.values())


def main():
    ks = load(sys.argv[1])
    agg = collections.OrderedDict()
    for k in ks:
        a = agg.setdefault(k["name"], [0, 0.0, 0.0])
        a[0] += 1
        a[1] += k.get("us", 0.0)
        a[2] += k.get("dram", 0.0)
    tot_us = sum(a[1] for a in agg.values())
    tot_b = sum(a[2] for a in agg.values())
    print("kernel\tlaunches\ttotal_ms\tavg_us\tshare\tdram_GB\tGB_per_s")
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n[:56]}\t{a[0]}\t{a[1] / 1e3:.3f}\t{a[1] / a[0]:.1f}\t{a[1] / tot_us:.3f}\t{a[2] / 1e9:.3f}\t"
              f"{(a[2] / 1e9) / (a[1] / 1e6) if a[1] else 0:.0f}")
    print(f"TOTAL\t{sum(a[0] for a in agg.values())}\t{tot_us / 1e3:.3f}\t\t1.000\t{tot_b / 1e9:.3f}\t")


if __name__ == "__main__":
    main()
