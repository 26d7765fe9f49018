"""TEST INFRASTRUCTURE: builds tests/_build/hostsim/libdie_hostsim.so -- die_b200/csrc compiled by g++ against
tests/hostsim/cuda_runtime.h, every CUDA thread a cooperative fiber on the CPU (see that header).

The kernel sources are used as they are, except for three pieces of syntax g++ cannot parse, rewritten on a copy:
  k<<<grid, block, smem, stream>>>(args)   ->  hostsim::launch(grid, block, smem, stream, k, args)
  extern __shared__ T name[];               ->  T* name = (T*)hostsim::dyn_smem();
  asm volatile("prefetch.global.L2 ...");   ->  (void)0;
The product (die_b200/) never imports this module and never loads the library it builds.
"""
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "die_b200", "csrc")
INCLUDE = os.path.join(ROOT, "include")
OUT_DIR = os.path.join(ROOT, "tests", "_build", "hostsim")
LIB = os.path.join(OUT_DIR, "libdie_hostsim.so")

_LAUNCH = re.compile(r"([A-Za-z_]\w*(?:<[^<>;()]*>)?)\s*<<<")
_DYN_SMEM = re.compile(r"extern\s+__shared__\s+([A-Za-z_][\w ]*?)\s+(\w+)\s*\[\s*\]\s*;")


def _split_top(s: str):
    parts, depth, cur = [], 0, []
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append("".join(cur).strip())
            cur = []
        else:
            cur.append(ch)
    parts.append("".join(cur).strip())
    return parts


def _rewrite_launches(src: str) -> str:
    out, pos = [], 0
    while True:
        m = _LAUNCH.search(src, pos)
        if m is None:
            out.append(src[pos:])
            return "".join(out)
        end = src.index(">>>", m.end())
        cfg = _split_top(src[m.end():end])
        if len(cfg) != 4:
            raise ValueError(f"launch configuration with {len(cfg)} parts: {src[m.start():end + 3]!r}")
        after = end + 3
        while src[after].isspace():
            after += 1
        if src[after] != "(":
            raise ValueError(f"no argument list after {src[m.start():end + 3]!r}")
        g, b, s, st = cfg
        out.append(src[pos:m.start()])
        out.append(f"hostsim::launch(dim3({g}), dim3({b}), (size_t)({s}), (cudaStream_t)({st}), {m.group(1)}, ")
        pos = after + 1


def _strip_asm(src: str) -> str:
    out, pos = [], 0
    while True:
        k = src.find("asm volatile(", pos)
        if k < 0:
            out.append(src[pos:])
            return "".join(out)
        i, depth, in_str = k + len("asm volatile("), 1, False
        while depth > 0:
            ch = src[i]
            if ch == '"' and src[i - 1] != "\\":
                in_str = not in_str
            elif not in_str:
                depth += (ch == "(") - (ch == ")")
            i += 1
        if "prefetch" not in src[k:i]:
            raise ValueError(f"inline PTX the emulator does not know: {src[k:i]!r}")
        while src[i] != ";":
            i += 1
        out.append(src[pos:k])
        out.append("(void)0;")
        pos = i + 1


def translate(src: str) -> str:
    if "DIE_HOSTSIM" not in src:            # (a file that names the macro keeps its PTX behind #if !defined(DIE_HOSTSIM))
        src = _strip_asm(src)
    src = _DYN_SMEM.sub(lambda m: f"{m.group(1)}* {m.group(2)} = ({m.group(1)}*)hostsim::dyn_smem();", src)
    src = _rewrite_launches(src)
    src = src.replace('#include "../../include/die_b200.h"', f'#include "{os.path.join(INCLUDE, "die_b200.h")}"')
    return src


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + \
        [os.path.join(HERE, f) for f in ("cuda_runtime.h", "hostsim.cpp", "build.py")] + [os.path.join(INCLUDE, "die_b200.h")]


def build(force: bool = False, csrc: str = CSRC, out_dir: str = OUT_DIR, extra_flags=()) -> str:
    lib = os.path.join(out_dir, "libdie_hostsim.so")
    if not force and csrc == CSRC and os.path.exists(lib) and os.path.getmtime(lib) >= max(os.path.getmtime(s) for s in _sources()):
        return lib
    os.makedirs(out_dir, exist_ok=True)
    for name in os.listdir(csrc):
        with open(os.path.join(csrc, name)) as f:
            text = translate(f.read())
        target = name[:-3] + ".cpp" if name.endswith(".cu") else name
        with open(os.path.join(out_dir, target), "w") as f:
            f.write(text)
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fno-strict-aliasing",
           "-shared", "-fPIC", "-DDIE_HOSTSIM", "-I", HERE, "-I", out_dir, *extra_flags, "-o", lib,
           os.path.join(out_dir, "die_api.cpp"), os.path.join(HERE, "hostsim.cpp")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + res.stdout + res.stderr[-6000:])
    return lib


if __name__ == "__main__":
    print(build(force=True))
