#!/bin/bash
# Evidence for the shipped default kernels (one GPU, after the plain run exited 0):
#   gpurun --timeout 1800 -- 'bash tools/gpu_profile.sh r02p'
#   1. the default bench line (no profiler)                         -> gpurun_out/<tag>_bench_default.json
#   2. ncu launch list of the same command (gpu__time_duration)     -> gpurun_out/<tag>_launches.csv
#   3. DRAM traffic of the four step kernels, batched + single      -> gpurun_out/<tag>_ncu_traffic_*.csv
#   4. ncu --set full + source of the four kernels (batch of 1024)  -> gpurun_out/prof_<tag>.ncu-rep, <tag>_ncu_full.txt
tag=${1:-prof}
out=gpurun_out
mkdir -p $out
K='gradient_forward_kernel|move_claim_kernel|field_step_kernel|agent_feed_kernel'
timeout 900 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_default.json 2> $out/${tag}_bench_default.err || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-small-env --no-f32 --steady= > $out/${tag}_launches.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    --kernel-name regex:"$K" --launch-skip 164 --launch-count 4 --csv --log-file $out/${tag}_ncu_traffic_batch4096.csv \
    python bench.py --workload batch256 --no-single-field --steps 3 --warmup 40 --no-e2e --no-cpu --no-f32 > $out/${tag}_ncu_traffic.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    --kernel-name regex:"$K" --launch-skip 164 --launch-count 4 --csv --log-file $out/${tag}_ncu_traffic_field4096.csv \
    python bench.py --workload field4096 --steps 3 --warmup 40 --no-e2e --no-cpu --no-f32 --steady= >> $out/${tag}_ncu_traffic.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name regex:"$K" \
    --launch-skip 164 --launch-count 4 -f -o $out/prof_${tag} \
    python bench.py --workload batch256 --batch 1024 --no-single-field --steps 3 --warmup 40 --no-e2e --no-cpu --no-f32 > $out/ncu_${tag}.log 2>&1
python tools/ncu_summary.py $out/prof_${tag}.ncu-rep > $out/${tag}_ncu_full.txt 2>&1
echo done
