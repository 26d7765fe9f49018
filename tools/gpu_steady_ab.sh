cd $GRAFT_REPO_ROOT
B="python bench.py --no-cpu --no-e2e --no-small-env --steady="
for variant in "--workload field4096 --warmup 3000 --steps 40" "--workload field4096 --warmup 3000 --steps 40 --tune fwd_lean=5" "--workload field4096 --warmup 3000 --steps 40 --tune fwd_min_blocks=5 --tune fwd_lean=0" "--workload field4096 --warmup 3000 --steps 40 --tune grad_f32=0" "--workload batch256 --no-single-field --warmup 600 --steps 40" ; do
  timeout 400 $B $variant > gpurun_out/r02i_tmp.json 2> gpurun_out/r02i_tmp.err
  python - "$variant" <<'PY' | tee -a gpurun_out/r02i_summary.txt
import json, sys
try:
    d = json.loads(open("gpurun_out/r02i_tmp.json").read().strip().splitlines()[-1])
    k = d["roofline"]["kernels"]
    print(f"{sys.argv[1]:90s} {d['ms_per_step']:.4f} ms " + " ".join(f"{n}={v['ms']:.4f}" for n, v in k.items()))
except Exception as exc:
    print(sys.argv[1], "FAILED", repr(exc), open("gpurun_out/r02i_tmp.err").read()[-500:])
PY
done
