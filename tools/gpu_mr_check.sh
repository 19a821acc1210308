#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_stress.py tests/test_gpu_multirank.py -x -q -m gpu 2>&1 | tail -5
