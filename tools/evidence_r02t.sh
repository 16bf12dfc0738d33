#!/bin/bash
# training-side glue kernels: backward parity tests, train-step timing, launch list of one training step
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_backward.py tests/test_gpu_train_extras.py tests/test_gpu_fullsize.py -q -s -p no:cacheprovider -k "not fullsize_vs_gpu and not caller" ) > gpurun_out/r02t_pytest.log 2>&1; tail -n 4 gpurun_out/r02t_pytest.log
for i in 1 2; do timeout 300 python tools/bench_train.py 8 256 256 2>&1 | tail -n 1 | cut -c1-330 >> gpurun_out/r02t_train.txt; done; cat gpurun_out/r02t_train.txt
BT_STEPS=1 BT_WARM=1 BIN_B200_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled --csv \
    --log-file gpurun_out/r02t_launches_train.csv python tools/bench_train.py 4 256 256 > gpurun_out/r02t_ncu_train.log 2>&1; tail -n 2 gpurun_out/r02t_ncu_train.log
