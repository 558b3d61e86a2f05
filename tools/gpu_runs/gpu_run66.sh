#!/bin/bash
timeout 300 python tools/attn_determinism.py 8 8 40 7488 10 2>&1 | tail -12
LDM_ATTN40=0 timeout 300 python tools/attn_determinism.py 8 8 40 7488 6 2>&1 | tail -8
timeout 300 python tools/attn_determinism.py 8 8 80 1872 10 2>&1 | tail -12
