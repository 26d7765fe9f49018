#!/bin/bash
# Early-regime and steady-state timing of the default kernels on one GPU (one summary line each):
#   gpurun --timeout 1500 -- 'bash tools/gpu_steady_ab.sh <tag> ["<extra bench args>" ...]'
tag=${1:-steady}; shift
out=gpurun_out
mkdir -p $out
B="python bench.py --no-cpu --no-e2e --no-small-env --no-f32 --steady="
[ $# -eq 0 ] && set -- ""
for extra in "$@"; do
for variant in "--workload batch256 --no-single-field --warmup 20 --steps 60" "--workload field4096 --warmup 20 --steps 60" "--workload field4096 --warmup 3000 --steps 40"; do
  timeout 400 $B $variant $extra > $out/${tag}_tmp.json 2> $out/${tag}_tmp.err
  python - "$variant $extra" $out/${tag}_tmp.json <<'PY' | tee -a $out/${tag}_summary.txt
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    k = d["roofline"]["kernels"]
    print(f"{sys.argv[1]:80s} {d['ms_per_step']:.4f} ms " + " ".join(f"{n}={v['ms']:.4f}" for n, v in k.items()))
except Exception as exc:
    print(sys.argv[1], "FAILED", repr(exc))
PY
done
done
