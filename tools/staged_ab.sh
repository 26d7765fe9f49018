#!/bin/bash
# First GPU call of the next round: does everything that was staged without a GPU hold on hardware, and what is it worth?
#   gpurun --timeout 1500 -- 'bash tools/staged_ab.sh r02a'          (append `ncu` for the traffic table + full capture)
# Writes gpurun_out/<tag>_*.  Every step runs under its own `timeout`; the mbarrier / cp.async.bulk kernel goes last
# (a protocol error traps after 2 s and poisons only its own process).
tag=${1:-staged}
out=gpurun_out
mkdir -p $out
B="python bench.py --no-cpu --no-e2e"

echo "== 1. the late-round-1 GPU tests (float32 vs float64 gradient cache, feed caps, unnormalised momentum)" | tee $out/${tag}_summary.txt
timeout 600 python -m pytest tests/test_gpu_zz_staged.py -q -rxXs 2>&1 | tail -15 | tee -a $out/${tag}_summary.txt

echo "== 2. A-B, batched 4096 x 256^2 and single 4096^2 (one JSON line each; compare ms_per_step and roofline.kernels)" | tee -a $out/${tag}_summary.txt
for variant in "" "--tune grad_f32=0" "--tune feed_min_blocks=4" "--tune feed_min_blocks=5"; do
    name=$(echo "base $variant" | tr -d '-' | tr ' =' '__')
    timeout 300 $B --steps 60 --warmup 20 $variant > $out/${tag}_ab_${name}.json 2> $out/${tag}_ab_${name}.err
    python - "$out/${tag}_ab_${name}.json" "$variant" <<'EOF' | tee -a $out/${tag}_summary.txt
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k = d["roofline"]["kernels"]
    single = list(d["also"].values())[0]
    print(f"{sys.argv[2] or 'base':45s} batched {d['ms_per_step']:.3f} ms  " + " ".join(f"{n}={v['ms']:.3f}" for n, v in k.items())
          + f" | single {single['ms_per_step']:.4f} ms " + " ".join(f"{n}={v['ms']:.4f}" for n, v in single['roofline']['kernels'].items()))
except Exception as exc:
    print(f"{sys.argv[2] or 'base':45s} FAILED: {exc!r}")
EOF
done

echo "== 3. host-buffer loop with 2 and 4 host threads" | tee -a $out/${tag}_summary.txt
for w in 2 4; do
    timeout 300 python bench.py --no-cpu --no-single-field --steps 20 --warmup 5 --e2e-workers $w > $out/${tag}_e2e_w$w.json 2> $out/${tag}_e2e_w$w.err
    python - "$out/${tag}_e2e_w$w.json" $w <<'EOF' | tee -a $out/${tag}_summary.txt
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(f"workers {sys.argv[2]}: one thread {d['e2e']['value']/1e6:.0f} M cell-updates/s, threads {d['e2e'].get('workers')}")
except Exception as exc:
    print(f"workers {sys.argv[2]}: FAILED {exc!r}")
EOF
done

echo "== 4. bulk-async field pass (field_impl=2): correctness first, in its own process, then timing" | tee -a $out/${tag}_summary.txt
DIE_B200_STAGED_BULK=1 timeout 300 python -m pytest tests/test_gpu_zz_staged.py -q -rxX -k bulk 2>&1 | tail -8 | tee -a $out/${tag}_summary.txt
if nvidia-smi > /dev/null 2>&1; then
    for variant in "--tune field_impl=2" "--tune field_impl=2 --tune grad_f32=0"; do
        name=$(echo "$variant" | tr -d '-' | tr ' =' '__')
        timeout 300 $B --steps 60 --warmup 20 $variant > $out/${tag}_ab_${name}.json 2> $out/${tag}_ab_${name}.err
        python - "$out/${tag}_ab_${name}.json" "$variant" <<'EOF' | tee -a $out/${tag}_summary.txt
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k = d["roofline"]["kernels"]
    single = list(d["also"].values())[0]
    print(f"{sys.argv[2]:45s} batched {d['ms_per_step']:.3f} ms  field_step={k['field_step']['ms']:.3f} | single {single['ms_per_step']:.4f} ms "
          f"field_step={single['roofline']['kernels']['field_step']['ms']:.4f}")
except Exception as exc:
    print(f"{sys.argv[2]:45s} FAILED: {exc!r}")
EOF
    done
fi
if [ "$2" = "ncu" ]; then
    echo "== 5. ncu: DRAM traffic of the step kernels (batched), full capture at steady state (single field)" | tee -a $out/${tag}_summary.txt
    timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        --kernel-name regex:"gradient_forward_kernel|move_claim_kernel|field_step_kernel|agent_feed_kernel" \
        --launch-skip 164 --launch-count 4 --csv --log-file $out/${tag}_ncu_traffic_batch4096.csv \
        python bench.py --workload batch256 --no-single-field --steps 3 --warmup 40 --no-e2e --no-cpu > $out/${tag}_ncu_traffic.log 2>&1
    timeout 900 bash tools/ncu_forward.sh ${tag}
    python tools/ncu_summary.py $out/prof_${tag}.ncu-rep > $out/${tag}_ncu_full_steady_state.txt 2>&1
    tail -5 $out/${tag}_ncu_traffic_batch4096.csv | cut -c1-300 | tee -a $out/${tag}_summary.txt
fi
echo "done" | tee -a $out/${tag}_summary.txt
