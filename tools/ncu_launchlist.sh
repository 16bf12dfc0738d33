export BIN_B200_GRAPH=0
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 802 -c 274 --csv --log-file gpurun_out/launches_window.csv python tools/run_window.py 2 > gpurun_out/ncu_launch.log 2>&1
tail -n 2 gpurun_out/ncu_launch.log
