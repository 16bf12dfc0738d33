#!/bin/bash
# final validation of the round-2 tree: full GPU suite, smoke, launch lists (window + train step), bench (both arms)
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q -s -p no:cacheprovider ) > gpurun_out/r02z_pytest.log 2>&1
tail -n 6 gpurun_out/r02z_pytest.log; grep -a "watchdog" gpurun_out/r02z_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02z_smoke.txt 2>&1; tail -n 1 gpurun_out/r02z_smoke.txt
export BIN_B200_GRAPH=0
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -s 227 -c 223 --csv \
    --log-file gpurun_out/r02z_launches_window.csv python tools/run_window.py 2 > gpurun_out/r02z_ncu_launch.log 2>&1
unset BIN_B200_GRAPH
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err
tail -c 900 gpurun_out/r02z_bench.json; tail -n 4 gpurun_out/r02z_bench.err
( time timeout 900 python bench.py --impl reference --steps 1 --warmup 1 ) > gpurun_out/r02z_bench_ref.json 2> gpurun_out/r02z_bench_ref.err
tail -c 600 gpurun_out/r02z_bench_ref.json
