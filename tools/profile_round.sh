#!/bin/bash
# Profiling pass for profiles/ (run under gpurun, one GPU):
#   (1) ncu --set full of ONE launch per hot kernel at bench shapes (tools/profile_kernels.py), ~1 min
#   (2) launch list with per-launch device time for ONE timed bench step at 8 new tokens (cold-cache, serialised: compare
#       SHARES), a few minutes: ncu costs ~45 ms per profiled launch with 35 GB resident, so only the NVTX range is profiled
# Usage: bash tools/profile_round.sh <tag> [--with-launch-list]
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:"gemm_bf16_tcgen05|gemm2_bf16_tcgen05|resample_u8|gemm_skinny|flash_attn|attn_decode|row_loraz|fbank|patchify_u8" -f -o gpurun_out/${TAG}_kernels \
    python tools/profile_kernels.py > gpurun_out/${TAG}_ncu.log 2>&1
if [ "$2" == "--with-launch-list" ]; then
  # only the timed step is profiled (NVTX range pushed by bench.py); everything before it runs at native speed
  timeout 420 ncu --nvtx --nvtx-include "crab_timed/" --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/${TAG}_launches.csv \
      python bench.py --steps 1 --warmup 3 --new-tokens 8 --no-cpu-baseline --profile-pass > gpurun_out/${TAG}_ncu_bench.log 2>&1
fi
ls -la gpurun_out/
