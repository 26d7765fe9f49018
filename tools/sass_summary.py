#!/usr/bin/env python
"""Memory-operation summary of the shipped kernels from the SASS of die_b200/libdie_sm100a.so (no GPU needed):
per kernel, how many LDG / STG of each width, shared-memory and TMA / mbarrier / cluster instructions there are.

    python tools/sass_summary.py [kernel-name regex] > profiles/<round>_sass_memory_ops.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ("LDG", "STG", "LD.", "ST.", "LDS", "STS", "ATOM", "RED", "UBLKCP", "UTMALDG", "SYNCS", "LDGSTS", "BAR", "UCGABAR",
        "MAPA", "CCTL", "DFMA", "DMUL", "DADD", "MUFU")


def main():
    pat = re.compile(sys.argv[1]) if len(sys.argv) > 1 else re.compile(
        r"gradient_forward_kernelILb1ELb0ELb0ELi4ELb1ELb1E|move_claim_kernelILb0ELb1E|field_step_kernelILi2ELi32ELi64ELi512ELb1ELb0ELb1E"
        r"|agent_feed_kernelILb0ELb0ELb1ELb0E|env_step_fused_kernelILi2ELi512ELb1ELb1E|brownian_forward")
    txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "die_b200", "libdie_sm100a.so")],
                         capture_output=True, text=True, check=True).stdout
    cur, counts, total = None, collections.defaultdict(collections.Counter), collections.Counter()
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1) if pat.search(m.group(1)) else None
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        total[cur] += 1
        if op.startswith(WANT):
            # keep the width / space qualifiers: LDG.E.64.CONSTANT -> LDG.E.64.CONSTANT
            counts[cur][op] += 1
    for fn in counts:
        demangled = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
        print(f"== {demangled[:150]}\n   {total[fn]} SASS instructions")
        for op, n in sorted(counts[fn].items()):
            print(f"   {n:5d}  {op}")


if __name__ == "__main__":
    main()
