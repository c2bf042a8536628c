#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_forward.py tests/test_checkpoint.py -m gpu -q -x -s 2>&1 | tail -25
