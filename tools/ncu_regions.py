#!/usr/bin/env python
"""Executed-instruction totals of the forward kernel grouped by source function (by line range).
   python tools/ncu_regions.py <report.ncu-rep> <slots>"""
import collections, csv, io, re, subprocess, sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_lines as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def func_ranges(path):
    """crude: map each line of a source file to the last seen function-like header"""
    out, cur = {}, 'file-scope'
    for ln, text in enumerate(open(path).read().splitlines(), 1):
        m = re.match(r'^(?:DIE_MATH_FN|__device__|__global__|template|static|inline).*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(', text)
        if m and not text.strip().startswith('//'):
            cur = m.group(1)
        m2 = re.match(r'^([a-z_0-9]+_kernel)\(', text)
        if m2:
            cur = m2.group(1)
        out[ln] = cur
    return out


def main():
    rep, slots = sys.argv[1], int(sys.argv[2])
    kre = sys.argv[3] if len(sys.argv) > 3 else 'gradient_forward'
    fsub = sys.argv[4] if len(sys.argv) > 4 else 'gradient_forward_kernelILb1'
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    h = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
    hdr = rows[h]
    ai, ii = hdr.index("Address"), hdr.index("Instructions Executed")
    lines = N.disasm_lines(fsub)
    ranges = {f: func_ranges(os.path.join(ROOT, 'die_b200', 'csrc', f)) for f in os.listdir(os.path.join(ROOT, 'die_b200', 'csrc'))}
    base, agg, ops, tot = None, collections.Counter(), collections.defaultdict(collections.Counter), 0
    for r in rows[h + 1:]:
        try:
            addr, n = int(r[ai], 16), int(r[ii])
        except (ValueError, IndexError):
            continue
        if base is None:
            base = addr
        key, ins = lines.get(addr - base, (None, ''))
        reg = 'unknown'
        if key:
            reg = f"{key[0]}:{ranges.get(key[0], {}).get(key[1], '?')}"
        agg[reg] += n
        tot += n
        op = re.sub(r'^@!?U?P\d+\s+', '', ins).split()[0].split('.')[0] if ins else '?'
        ops[reg][op] += n
    warps = slots / 32
    scale = 0.5 / warps          # the source page counts every instruction twice vs smsp__inst_executed
    print(f"instructions per slot-warp: {tot * scale:.0f}")
    for k, v in agg.most_common():
        print(f"{k:48s} {v * scale:7.1f}   " + ', '.join(f"{o}:{c * scale:.0f}" for o, c in ops[k].most_common(9)))


if __name__ == "__main__":
    main()
