#!/bin/bash
# GPU check + A-B timing of the cluster-fused environment step (die_env_fused.cuh) against the three kernels.
#   gpurun --timeout 1500 -- 'bash tools/gpu_fused_ab.sh r02b [ncu]'
tag=${1:-fused}
out=gpurun_out
mkdir -p $out
echo "== 1. parity: fused step vs three kernels vs oracle" | tee $out/${tag}_summary.txt
timeout 900 python -m pytest tests/test_gpu_fused_step.py -x -q 2>&1 | tail -12 | tee -a $out/${tag}_summary.txt
if ! nvidia-smi > /dev/null 2>&1; then echo "GPU lost" | tee -a $out/${tag}_summary.txt; exit 1; fi
echo "== 2. A-B timing (batched 4096 x 256^2 + single 4096^2)" | tee -a $out/${tag}_summary.txt
for variant in "" "--tune step_impl=0"; do
    name=$(echo "base $variant" | tr -d '-' | tr ' =' '__')
    timeout 400 python bench.py --no-cpu --no-e2e --steps 60 --warmup 20 $variant > $out/${tag}_ab_${name}.json 2> $out/${tag}_ab_${name}.err
    python - "$out/${tag}_ab_${name}.json" "$variant" <<'PY' | tee -a $out/${tag}_summary.txt
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k = d["roofline"]["kernels"]
    single = list(d["also"].values())[0]
    print(f"{sys.argv[2] or 'default':28s} batched {d['ms_per_step']:.3f} ms ({d['value']/1e9:.2f} G/s) " + " ".join(f"{n}={v['ms']:.3f}({v['frac']:.2f})" for n, v in k.items())
          + f" | single {single['ms_per_step']:.4f} ms " + " ".join(f"{n}={v['ms']:.4f}({v['frac']:.2f})" for n, v in single['roofline']['kernels'].items()))
except Exception as exc:
    print(f"{sys.argv[2] or 'default':28s} FAILED: {exc!r}")
PY
done
if [ "$2" = "ncu" ]; then
    echo "== 3. ncu: DRAM traffic + time of the default kernels (batched), full capture of the fused step" | tee -a $out/${tag}_summary.txt
    timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        --kernel-name regex:"gradient_forward_kernel|env_step_fused_kernel" \
        --launch-skip 80 --launch-count 2 --csv --log-file $out/${tag}_ncu_traffic_batch4096.csv \
        python bench.py --workload batch256 --no-single-field --steps 3 --warmup 40 --no-e2e --no-cpu > $out/${tag}_ncu_traffic.log 2>&1
    timeout 900 ncu --set full --import-source on --clock-control none \
        --kernel-name regex:'gradient_forward_kernel|env_step_fused_kernel' \
        --launch-skip 80 --launch-count 2 -f -o $out/prof_${tag}_batch \
        python bench.py --workload batch256 --batch 1024 --no-single-field --steps 3 --warmup 40 --no-e2e --no-cpu > $out/ncu_${tag}.log 2>&1
    python tools/ncu_summary.py $out/prof_${tag}_batch.ncu-rep > $out/${tag}_ncu_full_batch1024.txt 2>&1
    grep -v "^==" $out/${tag}_ncu_traffic_batch4096.csv | cut -d, -f5,13,15 | cut -c1-200 | tee -a $out/${tag}_summary.txt
fi
echo "done" | tee -a $out/${tag}_summary.txt
