#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline 2>/dev/null | wc -l
