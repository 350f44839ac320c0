"""profiles/r02_traffic.json from a whole-step ncu launch list summarised by scripts/launch_summary.py: DRAM bytes and time
per kernel family of bench.py's roofline object (bench.py reads `traffic_bytes_per_launch_mean` of the dominant family).

  python scripts/make_traffic_json.py profiles/r02_final_launches_b64.tsv > profiles/r02_traffic.json"""
import csv
import json
import sys


def family(k):
    if k.startswith("k_rb_chain"):
        return "conv_tc"
    if k.startswith("k_wgrad2_tc") or k.startswith("k_wgrad_reduce"):
        return "wgrad_tc"
    if k.startswith("k_im2col"):
        return "layout_tc"
    if k.startswith("k_col2im_g") or k.startswith("k_col2im<"):
        return "col2im"
    if k.startswith("k_col2im_coupling"):
        return "coupling_bwd" if k.rstrip(">").endswith("2") else "coupling_fwd"
    if k.startswith("k_hh_an_bwd"):
        return "hh_actnorm_bwd"
    if k.startswith("k_an_hh_fwd"):
        return "actnorm_hh_fwd"
    if k.startswith("k_squeeze"):
        return "squeeze"
    return None


def main():
    path = sys.argv[1]
    rows = list(csv.DictReader(open(path), delimiter="\t"))
    out = {"source": f"{path} (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, "
                     "scripts/step_probe.py cfg2 64 fp16x3: ONE whole training step of the round's final code, every launch of "
                     "every scale)"}
    agg = {}
    for r in rows:
        f = family(r["kernel"])
        if not f:
            continue
        a = agg.setdefault(f, {"kernels": [], "n": 0, "gb": 0.0, "ms": 0.0})
        a["kernels"].append(r["kernel"])
        a["n"] += int(r["launches"])
        a["gb"] += float(r["dram_GB"])
        a["ms"] += float(r["total_ms"])
    for f, a in agg.items():
        out[f] = {"kernels": a["kernels"], "launches_per_step": a["n"], "dram_GB_per_step": round(a["gb"], 3),
                  "ms_per_step_under_ncu": round(a["ms"], 3), "traffic_bytes_per_launch_mean": a["gb"] * 1e9 / a["n"]}
    t = [r for r in rows if r["kernel"] == "TOTAL"][0]
    out["step_total"] = {"launches": int(t["launches"]), "dram_GB": float(t["dram_GB"]), "ms_under_ncu": float(t["total_ms"])}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
