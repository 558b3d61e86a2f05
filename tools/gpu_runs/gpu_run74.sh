#!/bin/bash
timeout 900 python -m pytest tests/test_pipeline_gpu.py -q -x -k "main_ldm" 2>&1 | tail -15
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
