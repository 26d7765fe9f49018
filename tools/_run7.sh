out=gpurun_out; tag=r02zo
timeout 600 python -m pytest tests/test_gpu_committed_move.py tests/test_gpu_fused.py tests/test_gpu_philox_replay.py tests/test_gpu_graph_loop.py -x -q 2>&1 | tail -3
B="python bench.py --no-cpu --no-e2e --no-small-env --no-f32 --no-commit --no-single-field --warmup 20 --steps 60 --steady="
for round in 1 2; do
for v in "--tune cost_sqrt_near=0" "" "--tune cost_hint=0"; do
    timeout 500 $B $v > $out/${tag}_tmp.json 2> $out/${tag}_tmp.err
    python - "[$v]" $out/${tag}_tmp.json <<'PY' | tee -a $out/${tag}_cost_sqrt_ab.txt
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    k = d["roofline"]["kernels"]
    print(f"{sys.argv[1]:28s} {d['ms_per_step']:.4f} ms clk {d['clocks']['sm_mhz']} " + " ".join(f"{n}={v['ms']:.4f}" for n, v in k.items()))
except Exception as exc:
    print(sys.argv[1], "FAILED", repr(exc))
PY
done
done
