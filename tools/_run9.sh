out=gpurun_out; tag=r02zs
B="python bench.py --no-cpu --no-e2e --no-small-env --no-f32 --no-commit --warmup 20 --steps 60 --steady="
for v in "" "--tune field_tile=1" "--tune field_tile=2" ""; do
    timeout 500 $B $v > $out/${tag}_tmp.json 2> $out/${tag}_tmp.err
    python - "[$v]" $out/${tag}_tmp.json <<'PY' | tee -a $out/${tag}_field_tile_ab.txt
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    def show(name, e, clk=""):
        k = e["roofline"]["kernels"]
        print(f"{sys.argv[1]:24s} {name:16s} {e['ms_per_step']:.4f} ms {clk} " + " ".join(f"{n}={v['ms']:.4f}" for n, v in k.items()))
    show("batch4096x256^2", d, f"clk {d['clocks']['sm_mhz']}")
    for name, e in (d.get("also") or {}).items():
        show(name[-9:], e)
except Exception as exc:
    print(sys.argv[1], "FAILED", repr(exc))
PY
done
