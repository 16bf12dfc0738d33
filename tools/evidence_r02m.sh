#!/bin/bash
# Round-2 final-state bundle M: full GPU suite, smoke, launch list + ncu of the default kernels (QUAD), bench
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q -s -p no:cacheprovider ) > gpurun_out/r02m_pytest.log 2>&1
tail -n 6 gpurun_out/r02m_pytest.log; grep -a "watchdog\|\[bwd\|\[fullsize" gpurun_out/r02m_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02m_smoke.txt 2>&1; tail -n 1 gpurun_out/r02m_smoke.txt
export BIN_B200_GRAPH=0
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -s 227 -c 223 --csv \
    --log-file gpurun_out/r02m_launches_window.csv python tools/run_window.py 2 > gpurun_out/r02m_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:'conv_igemm_kernel<\(int\)32' -s 144 -c 3 -f -o gpurun_out/r02m_prof_conv_quad python tools/run_window.py 2 > gpurun_out/r02m_ncu_conv.log 2>&1
tail -n 2 gpurun_out/r02m_ncu_conv.log
unset BIN_B200_GRAPH
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r02m_bench.json 2> gpurun_out/r02m_bench.err
tail -c 900 gpurun_out/r02m_bench.json; tail -n 4 gpurun_out/r02m_bench.err
