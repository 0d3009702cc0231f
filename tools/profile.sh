#!/bin/bash
# Runs on the GPU box (under gpurun): kernel launch list of two bench steps + full ncu captures of the top kernels.
# Outputs land in gpurun_out/ ; tools/summarize_profiles.py (build box) turns them into the tracked files in profiles/.
set -u
mkdir -p gpurun_out
TAG=${1:-r01}
# 1) every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/launches_${TAG}.log 2>&1
# 2) the dominant kernel: FPN 3x3 256->256 implicit GEMM = the 21st igemm_persist_kernel launch of a forward pass
#    (igemm_persist_kernel<256, 4, true>, fused BatchNorm statistics), followed by the head 3x3 256->128; with source
timeout 600 ncu --set full --clock-control none --import-source on -k regex:igemm_persist_kernel --launch-skip 20 --launch-count 2 \
    -o gpurun_out/prof_igemm_${TAG} -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-graph > gpurun_out/prof_igemm_${TAG}.log 2>&1
# 3) the memory-bound head / loss kernels and one weight-gradient GEMM
timeout 900 ncu --set full --clock-control none -k 'regex:head_tail_fwd|head_tail_bwd_reduce|head_tail_bwd_apply|dbloss_reduce|dbloss_bwd|dbloss_select_pass2|wgrad_kernel|bn_bwd_apply|bn_bwd_reduce_fin|halo64_kernel' -c 26 \
    -o gpurun_out/prof_mem_${TAG} -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-graph > gpurun_out/prof_mem_${TAG}.log 2>&1
ls -la gpurun_out
