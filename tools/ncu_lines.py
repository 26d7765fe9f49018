#!/usr/bin/env python
"""Attribute executed SASS instructions of one kernel to CUDA source lines.

    python tools/ncu_lines.py <report.ncu-rep> <kernel regex> <mangled function substring> [topN]

Joins `ncu --page source --csv` (per-SASS-instruction executed counts and stall samples) with
`nvdisasm -g` line annotations of the cubin inside die_b200/libdie_sm100a.so, by code offset."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def disasm_lines(func_sub):
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "die_b200", "libdie_sm100a.so")],
                       cwd=d, check=True, capture_output=True)
        cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
        txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=d, check=True, capture_output=True, text=True).stdout
    out, cur, on = {}, None, False
    for line in txt.splitlines():
        if line.startswith(".text."):
            on = func_sub in line
            continue
        if not on:
            continue
        m = re.match(r'\s*//## File "(.*)", line (\d+)', line)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', line)
        if m:
            out[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return out


def main():
    rep, kre, fsub = sys.argv[1:4]
    topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    h = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
    hdr = rows[h]
    ai, ii, si = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
    lines = disasm_lines(fsub)
    base = None
    by_line, stall_by_line, tot, stot = collections.Counter(), collections.Counter(), 0, 0
    for r in rows[h + 1:]:
        try:
            addr, n = int(r[ai], 16), int(r[ii])
        except (ValueError, IndexError):
            continue
        if base is None:
            base = addr
        key = lines.get(addr - base, (None, ""))[0]
        by_line[key] += n
        stall_by_line[key] += int(r[si] or 0)
        tot += n
        stot += int(r[si] or 0)
    src_cache = {}
    print(f"total warp-instructions {tot}, stall samples {stot}")
    for key, n in by_line.most_common(topn):
        text = ""
        if key:
            f, ln = key
            if f not in src_cache:
                p = os.path.join(ROOT, "die_b200", "csrc", f)
                src_cache[f] = open(p).read().splitlines() if os.path.exists(p) else []
            if 0 < ln <= len(src_cache[f]):
                text = src_cache[f][ln - 1].strip()[:90]
        print(f"{n / tot * 100:6.2f}% instr {stall_by_line[key] / max(stot, 1) * 100:6.2f}% stall  {str(key):38s} {text}")


if __name__ == "__main__":
    main()
