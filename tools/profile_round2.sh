#!/bin/bash
# Round-2 profiling pass (run under gpurun, one GPU): (1) ncu --set full of the round-2 decode kernels, (2) the launch list of ONE
# timed bench step at 8 new tokens (NVTX range crab_timed; cold-cache, serialised: compare shares, not absolutes).
TAG=${1:-r03}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_skinny|decode_chain" -f -o gpurun_out/${TAG}_decode_kernels \
    python tools/profile_decode_r2.py > gpurun_out/${TAG}_ncu.log 2>&1
timeout 600 ncu --nvtx --nvtx-include "crab_timed/" --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --new-tokens 8 --no-cpu-baseline --legs "" --profile-pass > gpurun_out/${TAG}_ncu_bench.log 2>&1
ls -la gpurun_out/ | tail -5
