#!/bin/bash
mkdir -p gpurun_out
{
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python bench.py 2>/dev/null | cut -c1-400
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
} > gpurun_out/r2c30.txt 2>&1
