#!/bin/bash
# 1 GPU: compute-sanitizer over the kernel checks (tests/gpu_kernel_checks.py through pytest): memcheck on all of them,
# racecheck + synccheck on the kernels with hand-rolled synchronisation. Logs -> gpurun_out/sanitizer_*.log (copy the
# summaries to profiles/). The checks are small; under the sanitizer the whole script takes a few minutes.
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
K_ALL='not level0 and not long and not L0'
K_SYNC='groupnorm or splitk or attn_40_tail or attn_80 or attn_160 or gemm_plain_320 or gemm_conv3x3_L3 or gemm_qkv_40_vec or gemm_geglu or layernorm_320 or ccl or joint_hist'
timeout 900 $CS --tool memcheck --error-exitcode 3 --log-file gpurun_out/sanitizer_memcheck.log python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "$K_ALL" > gpurun_out/sanitizer_memcheck.pytest 2>&1
echo "memcheck rc=$?"; tail -2 gpurun_out/sanitizer_memcheck.pytest; tail -3 gpurun_out/sanitizer_memcheck.log
timeout 900 $CS --tool racecheck --racecheck-report all --error-exitcode 3 --log-file gpurun_out/sanitizer_racecheck.log python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "$K_SYNC" > gpurun_out/sanitizer_racecheck.pytest 2>&1
echo "racecheck rc=$?"; tail -2 gpurun_out/sanitizer_racecheck.pytest; tail -3 gpurun_out/sanitizer_racecheck.log
timeout 600 $CS --tool synccheck --error-exitcode 3 --log-file gpurun_out/sanitizer_synccheck.log python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "$K_SYNC" > gpurun_out/sanitizer_synccheck.pytest 2>&1
echo "synccheck rc=$?"; tail -2 gpurun_out/sanitizer_synccheck.pytest; tail -3 gpurun_out/sanitizer_synccheck.log
