#!/bin/bash
# Round-2 evidence bundle F (final state): full GPU suite, ncu captures for profiles/, bench (ours + reference arm).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r02f_gpu.txt
( time timeout 2400 python -m pytest tests -m gpu -q -s -p no:cacheprovider ) > gpurun_out/r02f_pytest.log 2>&1
tail -n 8 gpurun_out/r02f_pytest.log; grep -a "watchdog\|\[bwd\|\[fullsize\|\[weights" gpurun_out/r02f_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02f_smoke.txt 2>&1; tail -n 2 gpurun_out/r02f_smoke.txt
export BIN_B200_GRAPH=0
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -s 227 -c 223 --csv \
    --log-file gpurun_out/r02f_launches_window.csv python tools/run_window.py 2 > gpurun_out/r02f_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:'pack_frames_kernel|convlstm_kernel|conv_igemm_kernel<\(int\)16' -s 5 -c 5 -f -o gpurun_out/r02f_prof_k345 \
    python tools/run_window.py 2 > gpurun_out/r02f_ncu_k345.log 2>&1; tail -n 2 gpurun_out/r02f_ncu_k345.log
for pair in 0 1; do
  BIN_B200_PAIR=$pair timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
      -k regex:'conv_igemm_kernel<\(int\)32' -s 144 -c 3 -f -o gpurun_out/r02f_prof_conv_pair$pair python tools/run_window.py 2 > gpurun_out/r02f_ncu_conv$pair.log 2>&1
  tail -n 2 gpurun_out/r02f_ncu_conv$pair.log
done
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:'rdb_tail' -s 48 -c 1 -f -o gpurun_out/r02f_prof_tail python tools/run_window.py 2 > gpurun_out/r02f_ncu_tail.log 2>&1
unset BIN_B200_GRAPH
BT_STEPS=1 BT_WARM=1 BIN_B200_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled --csv \
    --log-file gpurun_out/r02f_launches_train.csv python tools/bench_train.py 4 256 256 > gpurun_out/r02f_ncu_train.log 2>&1; tail -n 2 gpurun_out/r02f_ncu_train.log
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err
tail -c 1200 gpurun_out/r02f_bench.json; tail -n 4 gpurun_out/r02f_bench.err
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/r02f_bench_ref.json 2> gpurun_out/r02f_bench_ref.err
cat gpurun_out/r02f_bench_ref.json; tail -n 4 gpurun_out/r02f_bench_ref.err
