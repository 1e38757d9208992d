#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_pytest.log 2>&1
echo "pytest exit $?"; tail -12 gpurun_out/t_pytest.log | cut -c1-300
