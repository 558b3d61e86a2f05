#!/bin/bash
timeout 100 python tools/profile_kernels.py --iters 20 --only logits_to_ids,segment_filter,decode_bitmap,joint_hist,ddim_step,layernorm,groupnorm 2>&1 | cut -c1-120
