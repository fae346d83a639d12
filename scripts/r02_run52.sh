#!/bin/bash
# round 2, GPU run 52 (2 GPUs): the multi-GPU pytest
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -4
