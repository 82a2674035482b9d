#!/bin/bash
o=gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -60 > $o/r2b_multi.log; tail -40 $o/r2b_multi.log
timeout 600 python tools/debug/rw_fp32_stats.py roughwell2d_mjhmc > $o/r2b_rwstats.log 2>&1; cat $o/r2b_rwstats.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_baseline_hp.py --deselect tests/test_gpu_multi.py 2>&1 | tail -30 > $o/r2b_pytest.log; tail -12 $o/r2b_pytest.log
