#!/bin/bash
timeout 600 python -m pytest tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -12
