#!/bin/bash
export DDK_TC=2
bash tools/tcr_trace.sh 2>&1 | grep -E "layer" | cut -c1-1100
