#!/bin/bash
timeout 900 python -m pytest tests/test_pipeline_gpu.py -m gpu -x -q -k "seg_encoder or seg_decoder" 2>&1 | tail -15
