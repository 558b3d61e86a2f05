#!/bin/bash
export LDM_GEMM_PAIR=0
for dbg in none notma nomma; do LDM_GEMM_DEBUG=$dbg python tools/gemm_dram_probe.py 2>&1 | tail -1; done
LDM_GEMM_STAGED=0 python tools/gemm_dram_probe.py 2>&1 | tail -1
