#!/bin/bash
# pair mode A-B on one GPU: parity tests, then early / steady-state timing of single fields with the mode off and on
#   gpurun --timeout 1500 -- 'bash tools/gpu_pair_ab.sh'
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_variants.py tests/test_gpu_graph_loop.py -x -q > $out/pair_tests.log 2>&1; echo "tests rc=$?" | tee $out/pair_summary.txt
tail -3 $out/pair_tests.log | tee -a $out/pair_summary.txt
B="python bench.py --no-cpu --no-e2e --no-small-env --no-f32 --steady="
for wl in ${SIZES:-4096 3072 8192}; do
for variant in "--warmup 20 --steps 60" "--warmup 3000 --steps 40"; do
for mode in 0 2; do
  timeout 500 $B --workload field4096 --field $wl $variant --tune pair_mode=$mode > $out/pair_tmp.json 2> $out/pair_tmp.err
  python - "field $wl $variant pair_mode=$mode" $out/pair_tmp.json <<'PY' | tee -a $out/pair_summary.txt
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    k = d["roofline"]["kernels"]
    print(f"{sys.argv[1]:60s} {d['ms_per_step']:.4f} ms {d['value']/1e9:.2f} G " + " ".join(f"{n}={v['ms']:.4f}" for n, v in k.items()))
except Exception as exc:
    print(sys.argv[1], "FAILED", repr(exc)); print(open(sys.argv[2].replace('.json', '.err')).read()[-600:])
PY
done; done; done
