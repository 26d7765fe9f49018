out=gpurun_out; tag=r02zt
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $out/${tag}_gpu_tests.txt 2>&1; echo "tests rc=$?"; tail -4 $out/${tag}_gpu_tests.txt | head -2
timeout 600 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_default_n1.json 2> $out/${tag}_bench_default_n1.err; echo "bench rc=$?"
B="python bench.py --no-cpu --no-e2e --no-small-env --no-f32 --no-commit --warmup 20 --steps 60 --steady="
for v in "--tune field_tile=1" "" "--tune field_tile=2"; do
    timeout 300 $B $v > $out/${tag}_tmp.json 2> $out/${tag}_tmp.err
    python - "[$v]" $out/${tag}_tmp.json <<'PY' | tee -a $out/${tag}_field_threads_ab.txt
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
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02zt_bench_default_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['clocks'], d['e2e']['value'], d['roofline']['frac'])
print({k: (v['ms'], v['frac']) for k, v in d['roofline']['kernels'].items()})
for k,v in d['also'].items():
    print(k, v.get('ms_per_step', v.get('ms_per_iter')), {kk: vv.get('ms_per_iter') for kk,vv in v.items() if isinstance(vv, dict) and 'ms_per_iter' in vv}, (v.get('steady_state') or {}).get('ms_per_step_in_the_40_steps_before_step'))
PY
