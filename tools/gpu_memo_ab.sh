#!/bin/bash
# the memoised forward (fwd_memo) against the LEAN forward on one GPU: parity tests, then timing early and at steady state
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_variants.py tests/test_gpu_philox_replay.py -x -q -k "memo or philox or pair" > $out/memo_tests.log 2>&1; echo "tests rc=$?" | tee $out/memo_summary.txt
tail -3 $out/memo_tests.log | tee -a $out/memo_summary.txt
B="python bench.py --no-cpu --no-e2e --no-small-env --no-f32 --steady="
for variant in "--workload batch256 --no-single-field --warmup 20 --steps 60" "--workload batch256 --no-single-field --warmup 600 --steps 40" "--workload field4096 --warmup 20 --steps 60" "--workload field4096 --warmup 3000 --steps 40"; do
for memo in 0 1 0 1; do
  timeout 500 $B $variant --tune fwd_memo=$memo > $out/memo_tmp.json 2> $out/memo_tmp.err
  python - "$variant fwd_memo=$memo" $out/memo_tmp.json <<'PY' | tee -a $out/memo_summary.txt
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    k = d["roofline"]["kernels"]
    print(f"{sys.argv[1]:84s} {d['ms_per_step']:.4f} ms clk {d['clocks']['sm_mhz']} " + " ".join(f"{n}={v['ms']:.4f}" for n, v in k.items()))
except Exception as exc:
    print(sys.argv[1], "FAILED", repr(exc)); print(open(sys.argv[2].replace('.json', '.err')).read()[-800:])
PY
done; done
