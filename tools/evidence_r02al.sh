#!/bin/bash
# x-stacked conv with the tile's 32 biases in registers: full GPU suite on the new library, then A/B against the previous one
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider ) > gpurun_out/r02al_pytest.log 2>&1
tail -n 5 gpurun_out/r02al_pytest.log
timeout 300 python tools/ab_conv.py prevlib > gpurun_out/r02al_biasreg.txt 2>&1
cut -c1-420 gpurun_out/r02al_biasreg.txt
