#!/bin/bash
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -k "dense or float32 or f32 or tc or pipelined" 2>&1 | tail -3
tools/variant_many.sh "main nosplit main nosplit" gauss100d_mjhmc_f32 pot100d_mjhmc_f32
