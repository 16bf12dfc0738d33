#!/bin/bash
# launch list of one training step (fwd+bwd) at batch 4 x 256x256 (steady state: skip packing + 2 warm-up steps)
export BIN_B200_GRAPH=0
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv \
    python tools/bench_train.py 4 256 256 > gpurun_out/ncu_train.log 2>&1
tail -n 2 gpurun_out/ncu_train.log
wc -l gpurun_out/launches_train.csv
