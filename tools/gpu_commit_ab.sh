#!/bin/bash
# A-B of the committed move (bench.py --fuse commit, DIE_FWD_COMMIT_MOVE) against the plain five-launch loop, on one box:
#   gpurun --timeout 1200 -- 'bash tools/gpu_commit_ab.sh <tag>'
tag=${1:-commitab}; out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_committed_move.py tests/test_gpu_fused.py tests/test_gpu_graph_loop.py -x -q \
    > $out/${tag}_tests.txt 2>&1; echo "tests rc=$?" | tee -a $out/${tag}_summary.txt; tail -3 $out/${tag}_tests.txt | tee -a $out/${tag}_summary.txt
B="python bench.py --no-cpu --no-e2e --no-small-env --no-f32 --warmup 20 --steps 60 --steady=3000"
for round in 1 2; do
for fuse in "" "--fuse commit"; do
    timeout 500 $B $fuse > $out/${tag}_tmp.json 2> $out/${tag}_tmp.err
    python - "[$fuse]" $out/${tag}_tmp.json <<'PY' | tee -a $out/${tag}_summary.txt
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    def show(name, e, clk=""):
        k = e["roofline"]["kernels"]
        print(f"{sys.argv[1]:16s} {name:12s} {e['ms_per_step']:.4f} ms {clk} " + " ".join(f"{n}={v['ms']:.4f}" for n, v in k.items()))
    show("batch4096x256^2", d, f"clk {d['clocks']['sm_mhz']} fused={d.get('fused_move')}")
    for name, e in (d.get("also") or {}).items():
        show(name[-9:], e)
        st = e.get("steady_state")
        if st:
            print(" " * 30, "steady:", json.dumps(st["ms_per_step_in_the_40_steps_before_step"]), json.dumps(st["kernel_ms_after_last"]))
except Exception as exc:
    print(sys.argv[1], "FAILED", repr(exc))
PY
    cp $out/${tag}_tmp.json "$out/${tag}_bench_r${round}_${fuse// /_}.json"
done
done
