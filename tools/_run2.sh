tag=r02zj; out=gpurun_out; mkdir -p $out
B="python bench.py --no-cpu --no-e2e --no-small-env --no-f32 --no-single-field --warmup 20 --steps 60 --steady="
for v in "" "--fuse commit" "--fuse commit --tune fwd_move_blocks=4"; do
  timeout 300 $B $v > $out/${tag}_tmp.json 2> $out/${tag}_tmp.err
  python - "[$v]" $out/${tag}_tmp.json <<'PY' | tee -a $out/${tag}_summary.txt
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    k = d["roofline"]["kernels"]
    print(f"{sys.argv[1]:50s} {d['ms_per_step']:.4f} ms clk {d['clocks']['sm_mhz']} " + " ".join(f"{n}={v['ms']:.4f}" for n, v in k.items()))
except Exception as exc:
    print(sys.argv[1], "FAILED", repr(exc))
PY
done
# ncu full of the committed forward kernel, batch of 1024 envs
ncu --set full --import-source on --clock-control none --kernel-name regex:'gradient_forward_kernel' \
    --launch-skip 30 --launch-count 1 -f -o $out/prof_${tag}_commit \
    python bench.py --batch 1024 --steps 3 --warmup 30 --no-e2e --no-cpu --no-small-env --no-f32 --no-single-field --fuse commit > $out/ncu_${tag}.log 2>&1
echo "ncu rc=$?"
python tools/ncu_summary.py $out/prof_${tag}_commit.ncu-rep | tee $out/${tag}_ncu_commit_forward.txt
