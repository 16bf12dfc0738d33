#!/bin/bash
# upper bound of fusing RDB convs 0..2: the same kernel with no input loads (and no stores) at all
mkdir -p gpurun_out
timeout 900 python tools/ab_conv.py noload > gpurun_out/r02aa_noload.txt 2>&1
cut -c1-330 gpurun_out/r02aa_noload.txt
