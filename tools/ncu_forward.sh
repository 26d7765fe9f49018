#!/bin/bash
# ncu --set full of the step kernels at steady state (single 4096^2 field), with source correlation.
# usage (on the GPU box): bash tools/ncu_forward.sh <tag> [extra bench args]
tag=$1; shift
ncu --set full --import-source on --clock-control none \
    --kernel-name regex:'gradient_forward_kernel|move_claim_kernel|field_step_kernel|agent_feed_kernel' \
    --launch-skip 164 --launch-count 4 -f -o gpurun_out/prof_${tag} \
    python bench.py --workload field4096 --steps 3 --warmup 40 --no-e2e --no-cpu "$@" > gpurun_out/ncu_${tag}.log 2>&1
echo "ncu rc=$?"
