#!/bin/bash
timeout 300 python tools/determinism_diag.py 2 16 24 40 2>&1 | tail -12
timeout 300 python tools/determinism_diag.py 8 48 156 6 2>&1 | tail -12
