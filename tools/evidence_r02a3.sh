#!/bin/bash
# debug: fp32-accurate mode at 720p (watchdog seen in the full-size test), then the remaining GPU tests
mkdir -p gpurun_out
BIN_B200_GRAPH=0 BIN_B200_DEBUG=16 timeout 300 python tools/run_window.py 2 720 1280 --fp32 > gpurun_out/r02a3_fp32_720p.txt 2>&1; tail -n 6 gpurun_out/r02a3_fp32_720p.txt
BIN_B200_GRAPH=0 timeout 300 python tools/run_window.py 2 256 256 --fp32 > gpurun_out/r02a3_fp32_256.txt 2>&1; tail -n 3 gpurun_out/r02a3_fp32_256.txt
timeout 300 python tools/run_window.py 2 720 1280 --fp32 > gpurun_out/r02a3_fp32_720p_graph.txt 2>&1; tail -n 4 gpurun_out/r02a3_fp32_720p_graph.txt
( time timeout 2400 python -m pytest tests -m gpu -q -s -p no:cacheprovider --deselect "tests/test_gpu_fullsize.py::test_window_fullsize_vs_gpu_oracle" --deselect tests/test_gpu_fullsize.py::test_reference_caller_sequence_720p ) > gpurun_out/r02a3_pytest.log 2>&1
tail -n 15 gpurun_out/r02a3_pytest.log
( time timeout 1200 python -m pytest tests/test_gpu_fullsize.py -q -s -p no:cacheprovider -k "fullsize_vs_gpu or caller_sequence" ) > gpurun_out/r02a3_pytest_fullsize.log 2>&1
tail -n 15 gpurun_out/r02a3_pytest_fullsize.log
