"""Top SASS instructions of an `ncu -i rep --page source --csv` export by warp-stall samples, with the dominant stall
reason of each (one table per profiled launch):

  ncu -i rep.ncu-rep --page source --csv > src.csv ; python scripts/ncu_hot_lines.py src.csv [N]"""
import csv
import sys


def val(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


def main():
    path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
    tables, hdr, cur, name = [], None, None, ""
    for r in csv.reader(open(path, errors="replace")):
        if r and r[0] == "Kernel Name":
            name = r[1]
            hdr = None
        elif r and r[0] == "Address":
            hdr, cur = r, []
            tables.append((name, hdr, cur))
        elif hdr is not None and len(r) == len(hdr):
            cur.append(r)
    for ti, (name, hdr, body) in enumerate(tables):
        k = hdr.index("# Samples")
        si = hdr.index("Source")
        stalls = [i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
        tot = sum(val(r[k]) for r in body) or 1.0
        by_reason = {hdr[i]: sum(val(r[i]) for r in body) for i in stalls}
        rs = sum(by_reason.values()) or 1.0
        print(f"== launch {ti}: {name[:60]}  {len(body)} SASS instructions, {tot:.0f} samples")
        print("   stall reasons (all warps): " + ", ".join(f"{n[6:]} {v / rs * 100:.0f}%" for n, v in sorted(by_reason.items(), key=lambda x: -x[1])[:8]))
        for r in sorted(body, key=lambda r: -val(r[k]))[:top]:
            dom = max(stalls, key=lambda i: val(r[i]))
            print(f"{val(r[k]) / tot * 100:6.2f}%  {r[si].strip()[:70]:<70s} {hdr[dom][6:]}")


if __name__ == "__main__":
    main()
