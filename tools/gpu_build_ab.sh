#!/bin/bash
# A-B of two BUILDS of the library on one GPU box: the in-tree one against tools/_ab_old.so (a copy of an earlier build)
#   gpurun --timeout 1500 -- 'bash tools/gpu_build_ab.sh <tag>'
tag=${1:-buildab}; out=gpurun_out; mkdir -p $out
B="python bench.py --no-cpu --no-e2e --no-small-env --no-f32 --steady="
cp die_b200/libdie_sm100a.so /tmp/_new.so
for round in 1 2; do
for which in new old; do
  if [ $which = old ]; then cp tools/_ab_old.so die_b200/libdie_sm100a.so; else cp /tmp/_new.so die_b200/libdie_sm100a.so; fi
  for variant in "--workload batch256 --no-single-field --warmup 20 --steps 60" "--workload field4096 --warmup 20 --steps 60"; do
    timeout 400 $B $variant > $out/${tag}_tmp.json 2> $out/${tag}_tmp.err
    python - "$which $variant" $out/${tag}_tmp.json <<'PY' | tee -a $out/${tag}_summary.txt
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    k = d["roofline"]["kernels"]
    print(f"{sys.argv[1]:70s} {d['ms_per_step']:.4f} ms clk {d['clocks']['sm_mhz']} " + " ".join(f"{n}={v['ms']:.4f}" for n, v in k.items()))
except Exception as exc:
    print(sys.argv[1], "FAILED", repr(exc))
PY
  done
done
done
cp /tmp/_new.so die_b200/libdie_sm100a.so
