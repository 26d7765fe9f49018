#!/bin/bash
# A-B timing of result-neutral variants on one GPU:  gpurun --timeout 1200 -- 'bash tools/gpu_ab.sh <tag> "<bench args A>" "<bench args B>" ...'
# every variant: python bench.py --no-cpu --no-e2e --steps 60 --warmup 20 <args>; one summary line each.
tag=$1; shift
out=gpurun_out
mkdir -p $out
: > $out/${tag}_summary.txt
n=0
for variant in "$@"; do
    n=$((n+1))
    timeout 400 python bench.py --no-cpu --no-e2e --steps 60 --warmup 20 $variant > $out/${tag}_ab_$n.json 2> $out/${tag}_ab_$n.err
    python - "$out/${tag}_ab_$n.json" "$variant" <<'PY' | tee -a $out/${tag}_summary.txt
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k = d["roofline"]["kernels"]
    line = f"{sys.argv[2] or 'default':44s} {d['config']['workload'][9:22]} {d['ms_per_step']:.3f} ms ({d['value']/1e9:.2f} G/s) " + " ".join(f"{n}={v['ms']:.3f}({v['frac']:.2f})" for n, v in k.items())
    if d.get("also"):
        single = list(d["also"].values())[0]
        line += f" | single {single['ms_per_step']:.4f} ms " + " ".join(f"{n}={v['ms']:.4f}({v['frac']:.2f})" for n, v in single['roofline']['kernels'].items())
    print(line)
except Exception as exc:
    print(f"{sys.argv[2] or 'default':44s} FAILED: {exc!r}")
PY
done
