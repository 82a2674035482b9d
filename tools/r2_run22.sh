#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x -k "dense or pot or ProductOfT or Gaussian or gauss or baseline or screen" 2>&1 | tail -5
tools/variant_many.sh "main denseold main denseold" gauss100d_mjhmc pot100d_mjhmc
