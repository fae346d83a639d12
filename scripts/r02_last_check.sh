#!/bin/bash
# round 2, last GPU seconds: the drop-in CLI against the real library after the shim's group formatter
timeout 200 python -m pytest tests/test_gpu_cli.py -x -q 2>&1 | tail -2
