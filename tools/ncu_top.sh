#!/bin/bash
# ncu captures for the round summary (run under gpurun, 1 GPU). Outputs -> gpurun_out/
set -x
export BIN_B200_GRAPH=0
mkdir -p gpurun_out
# (1) launch list of ONE steady-state window: skip weight packing (4 backbones x 132) + window 0 (274)
ncu --metrics gpu__time_duration.sum --clock-control none -s 802 -c 274 --csv \
    --log-file gpurun_out/launches_window.csv python tools/run_window.py 2 > gpurun_out/ncu_launch.log 2>&1
# (2) full capture of the dominant kernel (x-stacked RDB conv), 3 launches from the steady state
# conv launches only: window 0 = 330; window 1 stage 1 (5 batched calls): idx 2+5*i+c -> RDB 5 = 357..361
ncu --set full --clock-control none --import-source on -k regex:conv_igemm_kernel -s 357 -c 5 \
    -o gpurun_out/prof_rdb5 python tools/run_window.py 2 > gpurun_out/ncu_full.log 2>&1
# (3) tail of the same stage: GFF.0, GFF.1, UPNet.0, UPNet.2 (conv idx 62..65 of the stage)
ncu --set full --clock-control none --import-source on -k regex:conv_igemm_kernel -s 392 -c 4 \
    -o gpurun_out/prof_tail python tools/run_window.py 2 > gpurun_out/ncu_full2.log 2>&1
tail -n 3 gpurun_out/ncu_launch.log gpurun_out/ncu_full.log gpurun_out/ncu_full2.log
ls -la gpurun_out
