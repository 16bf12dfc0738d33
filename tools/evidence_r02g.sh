#!/bin/bash
# Round-2 evidence bundle G: zigzag tile order A/B (whole window, graph replay) + bit-identity of every switch combination
mkdir -p gpurun_out
for z in 0 1 0 1; do BIN_B200_ZIGZAG=$z timeout 300 python tools/run_window.py 8 --graph 2>&1 | tail -n 4 | sed "s/^/zigzag=$z /" >> gpurun_out/r02g_window_zigzag.txt; done; cat gpurun_out/r02g_window_zigzag.txt
( time timeout 900 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -k "cta_pair" ) > gpurun_out/r02g_pytest_switches.log 2>&1; tail -n 4 gpurun_out/r02g_pytest_switches.log
export BIN_B200_GRAPH=0
for z in 0 1; do
  BIN_B200_ZIGZAG=$z timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none --kernel-name-base demangled \
    -k regex:'rdb_tail|conv_igemm_kernel<\(int\)32' -s 192 -c 8 --csv --log-file gpurun_out/r02g_dram_zigzag$z.csv python tools/run_window.py 2 > gpurun_out/r02g_ncu_z$z.log 2>&1
done
unset BIN_B200_GRAPH
tail -n 3 gpurun_out/r02g_ncu_z1.log
