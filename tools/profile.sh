#!/bin/bash
# Runs on the GPU box (under gpurun): kernel launch list of two bench steps + one full ncu capture of the top kernels.
# Outputs land in gpurun_out/ ; summaries are copied to profiles/ by tools/summarize_profiles.py on the build box.
set -u
mkdir -p gpurun_out
TAG=${1:-r01}
# 1) every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_${TAG}.log 2>&1
# 2) full capture: the FPN 256->256 3x3 implicit GEMM (dominant kernel), one N=64 GEMM, the head tail and the loss kernels
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:igemm_kernel<256|head_tail_fwd|head_tail_bwd_apply|dbloss_reduce|dbloss_bwd' -c 8 \
    -o gpurun_out/prof_${TAG} -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/prof_${TAG}.log 2>&1
ls -la gpurun_out
