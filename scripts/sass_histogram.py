"""Static SASS mnemonic histogram of the built library (cuobjdump -sass): per kernel and in total, the opcodes that prove
the tcgen05 / TMA / TMEM path (UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA load / store, LDTM / STTM = tcgen05.ld / st).

  python scripts/sass_histogram.py > profiles/r02_sass_histogram.json"""
import collections
import json
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "invertiblenetworks.jl_b200", "libinb200.so")
KEYS = ["UTCHMMA", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "UTCBAR", "SYNCS", "ELECT", "FFMA", "DFMA", "REDUX", "SHFL",
        "MUFU", "LDS", "STS", "LDG", "STG", "PRMT"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    per, tot, cur, i, ninstr = {}, collections.Counter(), None, 0, 0
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = per.setdefault(names[i], collections.Counter())
            i += 1
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["instructions"] += 1
            ninstr += 1
            for k in KEYS:
                if op.startswith(k):
                    cur[k] += 1
                    tot[k] += 1
    out = {"source": "cuobjdump -sass invertiblenetworks.jl_b200/libinb200.so (sm_100a); mnemonic counts per kernel (static instructions)",
           "library_totals": {k: tot[k] for k in KEYS}, "functions": len(per), "instructions": ninstr,
           "kernels": {n: dict(c) for n, c in per.items() if any(c[k] for k in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM"))}}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
