"""Registers / spills / static shared memory of every kernel in libdie_sm100a.so, from `ptxas -v` (no GPU needed).

    python tools/ptxas_table.py [pattern]

One line per kernel instantiation: what to look at before spending GPU time on a variant (a register cap that
spills, an instantiation that lost an occupancy step)."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from die_b200._build import NVCC_FLAGS, _nvcc, CSRC      # noqa: E402


def main():
    pat = re.compile(sys.argv[1]) if len(sys.argv) > 1 else None
    flags = [f for f in NVCC_FLAGS if f not in ("-shared", "-Xcompiler", "-fPIC", "-lineinfo")]
    res = subprocess.run([_nvcc(), *flags, "-Xptxas=-v", "-c", "-o", os.devnull, os.path.join(CSRC, "die_api.cu")],
                         capture_output=True, text=True)
    if res.returncode != 0:
        sys.exit(res.stderr)
    rows, name = [], None
    for line in res.stderr.splitlines():
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            name = m.group(1)
            spill = None
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m and name:
            spill = (int(m.group(1)), int(m.group(2)), int(m.group(3)))
            continue
        m = re.search(r"Used (\d+) registers(?:, used \d+ barriers)?(?:, (\d+) bytes smem)?", line)
        if m and name:
            rows.append((name, int(m.group(1)), spill, int(m.group(2) or 0)))
            name = None
    names = subprocess.run(["cu++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
    print(f"{'registers':>9} {'CTAs/SM@256':>11} {'stack':>6} {'spill st/ld':>12} {'smem':>6}  kernel")
    for (mangled, regs, spill, smem), full in zip(rows, names):
        cut = full.find(">(")
        short = full[:cut + 1] if cut >= 0 else full.split("(")[0]
        short = short.replace("void die::", "").replace("void ", "").replace("(bool)1", "1").replace("(bool)0", "0").replace("(int)", "")
        if pat and not pat.search(short):
            continue
        ctas = min(8, 65536 // (256 * ((regs + 7) // 8 * 8)))
        st = spill or (0, 0, 0)
        print(f"{regs:>9} {ctas:>11} {st[0]:>6} {st[1]:>5}/{st[2]:<6} {smem:>6}  {short}")


if __name__ == "__main__":
    main()
