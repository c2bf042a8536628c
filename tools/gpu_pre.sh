#!/bin/bash
timeout 600 python -m pytest tests/test_preprocess.py tests/test_host.py -q -x 2>&1 | tail -8
