#!/bin/bash
# shared-memory-state kernel: parity against the register kernel, occupancy variants on the Funnel / Gaussian points
o=gpurun_out
tag=${1:-r2q}
timeout 900 python -m pytest tests/test_gpu_screen.py -q -x --timeout 300 2>&1 | tail -8
echo "--- register kernel"; MJHMC_B200_REGISTER_STATE=1 tools/variant_many.sh "main" funnel10d_cthmc funnel10d_cthmc_ess
echo "--- stash variants"; tools/variant_many.sh "main st4 st6" funnel10d_cthmc funnel10d_cthmc_ess
tools/variant_many.sh "main" gauss100d_diag_mjhmc roughwell2d_mjhmc
