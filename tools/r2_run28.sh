#!/bin/bash
tools/probe/tma_probe 0 64; tools/probe/tma_probe 0 65; tools/probe/tma_probe 0 68
tools/r2_run26.sh
