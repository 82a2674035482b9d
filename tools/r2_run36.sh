#!/bin/bash
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 400 -x -k "tma_boxes and ProductOfT and MarkovJumpHMC or job_ranges and Gaussian or dense_float32_tcgen05 and 100 and MarkovJumpHMC" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|Error" | head -12
