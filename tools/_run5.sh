out=gpurun_out; tag=r02zm
timeout 900 python -m pytest tests/test_gpu_committed_move.py tests/test_gpu_fused.py tests/test_gpu_graph_loop.py tests/test_gpu_philox_replay.py tests/test_gpu_variants.py -x -q 2>&1 | tail -4
B="python bench.py --no-cpu --no-e2e --no-small-env --no-f32 --no-commit --warmup 20 --steps 60 --steady=3000"
for round in 1 2; do
for v in "--tune cost_hint=0" ""; do
    timeout 500 $B $v > $out/${tag}_tmp.json 2> $out/${tag}_tmp.err
    python - "[$v]" $out/${tag}_tmp.json <<'PY' | tee -a $out/${tag}_cost_hint_ab.txt
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    def show(name, e, clk=""):
        k = e["roofline"]["kernels"]
        print(f"{sys.argv[1]:22s} {name:16s} {e['ms_per_step']:.4f} ms {clk} " + " ".join(f"{n}={v['ms']:.4f}" for n, v in k.items()))
    show("batch4096x256^2", d, f"clk {d['clocks']['sm_mhz']}")
    for name, e in (d.get("also") or {}).items():
        show(name[-9:], e)
        st = e.get("steady_state")
        if st:
            print(" " * 30, "steady:", json.dumps(st["ms_per_step_in_the_40_steps_before_step"]), json.dumps(st["kernel_ms_after_last"]))
except Exception as exc:
    print(sys.argv[1], "FAILED", repr(exc))
PY
done
done
