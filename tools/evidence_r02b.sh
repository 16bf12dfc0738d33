#!/bin/bash
# Round-2 evidence bundle B (run under gpurun, 1 GPU): CTA-pair kernels -- numerics, A/B, ncu before/after, launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r02b_gpu.txt
# 0. the whole GPU suite in one process (pair kernels off by default)
( time timeout 2400 python -m pytest tests -m gpu -q -s -p no:cacheprovider ) > gpurun_out/r02b_pytest.log 2>&1
tail -n 12 gpurun_out/r02b_pytest.log; grep -a "watchdog\|\[bwd\|\[fullsize\|\[weights" gpurun_out/r02b_pytest.log
# 1. numerics of everything that touches the pair kernels, with the pair kernels on
( time BIN_B200_PAIR=1 timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_conv_fuzz.py tests/test_gpu_streaming.py -q -p no:cacheprovider ) > gpurun_out/r02b_pytest_pair.log 2>&1
tail -n 8 gpurun_out/r02b_pytest_pair.log
# 2. A/B timings in separate processes
timeout 600 python tools/ab_pair.py > gpurun_out/r02b_ab_pair.txt 2>&1; cat gpurun_out/r02b_ab_pair.txt
for pair in 0 1; do BIN_B200_PAIR=$pair timeout 300 python tools/fusion_bound.py >> gpurun_out/r02b_fusion_bound.txt 2>&1; done; cat gpurun_out/r02b_fusion_bound.txt
for pair in 0 1 0 1; do BIN_B200_PAIR=$pair timeout 300 python tools/run_window.py 6 --graph 2>&1 | tail -n 4 | sed "s/^/pair=$pair /" >> gpurun_out/r02b_window_ab.txt; done; cat gpurun_out/r02b_window_ab.txt
# 3. ncu: launch list of one steady-state window (pair on), then --set full of the two RDB kernels, pair off / on
export BIN_B200_GRAPH=0
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -s 227 -c 223 --csv \
    --log-file gpurun_out/r02b_launches_window.csv python tools/run_window.py 2 > gpurun_out/r02b_ncu_launch.log 2>&1
tail -n 2 gpurun_out/r02b_ncu_launch.log
for pair in 0 1; do
  BIN_B200_PAIR=$pair timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
      -k regex:'rdb_tail' -s 48 -c 1 -f -o gpurun_out/r02b_prof_tail_pair$pair python tools/run_window.py 2 > gpurun_out/r02b_ncu_tail$pair.log 2>&1
  BIN_B200_PAIR=$pair timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
      -k regex:'conv_igemm_kernel<32' -s 144 -c 3 -f -o gpurun_out/r02b_prof_conv_pair$pair python tools/run_window.py 2 > gpurun_out/r02b_ncu_conv$pair.log 2>&1
  tail -n 2 gpurun_out/r02b_ncu_tail$pair.log gpurun_out/r02b_ncu_conv$pair.log
done
unset BIN_B200_GRAPH
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
tail -c 2500 gpurun_out/r02b_bench.json; tail -n 4 gpurun_out/r02b_bench.err
ls -la gpurun_out | tail -20
