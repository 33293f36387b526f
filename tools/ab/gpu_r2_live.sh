#!/bin/bash
timeout 600 python -m pytest tests/test_live_gpu.py tests/test_dropin_gpu.py -m gpu -x -q 2>&1 | tail -25
