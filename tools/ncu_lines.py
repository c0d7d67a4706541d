#!/usr/bin/env python
"""Attribute an ncu capture's per-instruction samples / executed counts to CUDA source lines.
ncu's CSV source page carries SASS only; the line table comes from nvdisasm -g on the same cubin, matched by
instruction order.   Usage: ncu_lines.py prof.ncu-rep <kernel-substring> [top]"""
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gnn_jet_autoencoder_b200", "libgnnjet_b200.so")


def line_table(kernel_sub):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, capture_output=True)
    for f in sorted(os.listdir(tmp)):
        if not f.endswith(".cubin"):
            continue
        sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if kernel_sub not in sass:
            continue
        out, cur, inside = [], None, False
        for ln in sass.splitlines():
            if ln.startswith(".text."):
                inside = kernel_sub in ln
                continue
            if not inside:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
                out.append((cur, ln.strip()))
        if out:
            return out
    raise SystemExit("kernel not found in any cubin")


def main():
    rep, ksub = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    table = line_table(ksub)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = next(r for r in rows if r and r[0] == "Address")
    data = [dict(zip(hdr, r)) for r in rows if len(r) == len(hdr) and r[0].startswith("0x")]
    if len(data) != len(table):
        print(f"warning: {len(data)} profiled instructions vs {len(table)} disassembled", file=sys.stderr)
    f = lambda x: float(x.replace(",", "")) if x else 0.0
    agg = {}
    for (loc, _), d in zip(table, data):
        a = agg.setdefault(loc, [0.0, 0.0, 0.0])
        a[0] += f(d["# Samples"]); a[1] += f(d["Instructions Executed"]); a[2] += f(d.get("stall_long_sb", "0"))
    ts, ti = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
    src = {}
    print(f"total samples {ts:.0f}, warp instructions {ti:.0f}")
    for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        text = ""
        if loc:
            path = os.path.join(ROOT, "gnn_jet_autoencoder_b200", "csrc", loc[0])
            if path not in src and os.path.exists(path):
                src[path] = open(path).read().splitlines()
            if path in src and loc[1] <= len(src[path]):
                text = src[path][loc[1] - 1].strip()[:95]
        print(f"{100*a[0]/ts:5.1f}% smp {100*a[1]/ti:5.1f}% inst {100*a[2]/max(ts,1):5.1f}% longsb  {loc[0] if loc else '?'}:{loc[1] if loc else 0:<4d} {text}")


if __name__ == "__main__":
    main()
