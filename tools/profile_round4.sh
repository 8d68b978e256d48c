#!/bin/bash
# Final round-2 profiling pass (run under gpurun, one GPU): (1) ncu --set full of the decode linears in their final form,
# (2) the launch list of ONE timed bench step at 8 new tokens (NVTX range crab_timed; cold-cache, serialised: compare shares).
TAG=${1:-r04}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_skinny" -f -o gpurun_out/${TAG}_decode_kernels \
    python tools/profile_decode_r4.py > gpurun_out/${TAG}_ncu.log 2>&1
timeout 900 ncu --nvtx --nvtx-include "crab_timed/" --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --new-tokens 8 --no-cpu-baseline --legs "" --profile-pass > gpurun_out/${TAG}_ncu_bench.log 2>&1
ls -la gpurun_out/ | tail -5
