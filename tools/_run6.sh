out=gpurun_out; tag=r02zp
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $out/${tag}_gpu_tests.txt 2>&1; echo "tests rc=$?"; tail -4 $out/${tag}_gpu_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
bash tools/gpu_profile.sh $tag 2>&1 | tail -3
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $out/${tag}_bench_reference_n1.json 2>&1
python tools/sass_summary.py > $out/${tag}_sass_memory_ops.txt 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02zp_bench_default.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['clocks'], d['e2e']['value'], d['roofline']['frac'])
print({k: (v['ms'], v['frac']) for k, v in d['roofline']['kernels'].items()})
for k,v in d['also'].items():
    print(k, v.get('ms_per_step', v.get('ms_per_iter')), {kk: vv.get('ms_per_iter') for kk,vv in v.items() if isinstance(vv, dict) and 'ms_per_iter' in vv}, (v.get('steady_state') or {}).get('ms_per_step_in_the_40_steps_before_step'))
PY
