out=gpurun_out; tag=r02zk
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $out/${tag}_gpu_tests.txt 2>&1; echo "tests rc=$?"; tail -4 $out/${tag}_gpu_tests.txt
( time timeout 900 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_default_n1.json 2> $out/${tag}_bench_default_n1.err ); echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02zk_bench_default_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['clocks'], d['e2e']['value'])
for k,v in d['also'].items():
    print(k, v.get('ms_per_step', v.get('ms_per_iter')), {kk: vv.get('ms_per_iter') for kk,vv in v.items() if isinstance(vv, dict) and 'ms_per_iter' in vv})
PY
