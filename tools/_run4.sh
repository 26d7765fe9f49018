out=gpurun_out; tag=r02zl
timeout 600 python -m pytest tests/test_gpu_streams_examples.py tests/test_gpu_jones.py -x -q 2>&1 | tail -5
timeout 900 python bench.py --steps 20 --warmup 5 --no-small-env --no-f32 --no-single-field --no-commit > $out/${tag}_bench_e2e.json 2> $out/${tag}_bench_e2e.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02zl_bench_e2e.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['clocks'])
print(json.dumps(d['e2e'])[:1500])
PY
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $out/${tag}_bench_reference.json 2>&1; tail -c 600 $out/${tag}_bench_reference.json
