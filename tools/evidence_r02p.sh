#!/bin/bash
# final validation after restricting QUAD to epilogues without a residual: train A/B, full GPU suite, bench
mkdir -p gpurun_out
for q in 0 1 0 1; do BIN_B200_QUAD=$q timeout 300 python tools/bench_train.py 8 256 256 2>&1 | tail -n 1 | cut -c1-200 | sed "s/^/quad=$q /" >> gpurun_out/r02p_train.txt; done; cat gpurun_out/r02p_train.txt
( time timeout 2400 python -m pytest tests -m gpu -q -s -p no:cacheprovider ) > gpurun_out/r02p_pytest.log 2>&1
tail -n 6 gpurun_out/r02p_pytest.log; grep -a "watchdog" gpurun_out/r02p_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02p_smoke.txt 2>&1; tail -n 1 gpurun_out/r02p_smoke.txt
export BIN_B200_GRAPH=0
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -s 227 -c 223 --csv \
    --log-file gpurun_out/r02p_launches_window.csv python tools/run_window.py 2 > gpurun_out/r02p_ncu_launch.log 2>&1
unset BIN_B200_GRAPH
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r02p_bench.json 2> gpurun_out/r02p_bench.err
tail -c 700 gpurun_out/r02p_bench.json; tail -n 4 gpurun_out/r02p_bench.err
