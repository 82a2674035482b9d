#!/bin/bash
o=gpurun_out
tools/probe/fp64_probe
echo "--- variants"; tools/variant_many.sh "main st3 es_reg es_st4" funnel10d_cthmc funnel10d_cthmc_ess
