# launch list of one steady-state window + one --set full capture of the fused RDB tail kernel
export BIN_B200_GRAPH=0
SKIP=${SKIP:-754}     # 528 weight-pack launches + the 226 launches of window 0
COUNT=${COUNT:-226}   # 4 batched backbone stages x (1 pack + 54 conv) + 6 ConvLSTM
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c $COUNT --csv --log-file gpurun_out/launches_window.csv python tools/run_window.py 2 > gpurun_out/ncu_launch.log 2>&1
tail -n 2 gpurun_out/ncu_launch.log
ncu --set full --clock-control none --import-source on -k regex:rdb_tail_kernel -s 48 -c 2 -o gpurun_out/prof_rdb_tail -f python tools/run_window.py 2 > gpurun_out/ncu_full_tail.log 2>&1
tail -n 3 gpurun_out/ncu_full_tail.log
