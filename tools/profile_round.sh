#!/bin/bash
# Profiling pass for profiles/: (1) launch list with per-launch device time, (2) ncu --set full of the top kernels.
# Usage (under gpurun): bash tools/profile_round.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
ARGS="--steps 1 --warmup 3 --new-tokens 6 --no-cpu-baseline"
# (1) all launches of a short run (3 warm-up + 1 timed + e2e + eager split); post-processed by tools/summarize_launches.py
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py $ARGS > gpurun_out/${TAG}_ncu_bench.log 2>&1
# (2) full sets, 2 launches each, skipping the warm-up passes
for K in gemm_bf16_tcgen05_kernel gemm_skinny_tcgen05_kernel attn_decode_kernel flash_attn_kernel row_loraz_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 600 -c 2 -f -o gpurun_out/${TAG}_$K \
      python bench.py $ARGS > gpurun_out/${TAG}_ncu_$K.log 2>&1
done
ls -la gpurun_out/
