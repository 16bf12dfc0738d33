#!/bin/bash
# Round-1 evidence bundle (run under gpurun, 1 GPU): role timelines, training launch list, wgrad ncu capture, memcheck.
mkdir -p gpurun_out
python tools/timeline.py 96 > gpurun_out/timeline_rdbconv0.txt 2>&1
python tools/timeline2.py final > gpurun_out/timeline_final.txt 2>&1
python tools/timeline2.py lff > gpurun_out/timeline_lff.txt 2>&1
export BIN_B200_GRAPH=0
BT_STEPS=1 BT_WARM=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv \
    python tools/bench_train.py 4 256 256 > gpurun_out/ncu_train.log 2>&1
BT_STEPS=1 BT_WARM=1 ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 300 -c 2 -o gpurun_out/prof_wgrad \
    python tools/bench_train.py 4 256 256 > gpurun_out/ncu_wgrad.log 2>&1
timeout 300 compute-sanitizer --tool memcheck --print-limit 10 python tools/bench_train.py 1 64 64 > gpurun_out/memcheck_train.txt 2>&1
tail -n 3 gpurun_out/memcheck_train.txt
ls -la gpurun_out | tail -12
