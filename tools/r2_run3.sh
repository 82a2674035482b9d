#!/bin/bash
o=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x --timeout 60 -k "dense_float32" 2>&1 | tail -40 > $o/r2c_tc.log; tail -40 $o/r2c_tc.log
