#!/bin/bash
# Round-2 evidence bundle A (run under gpurun, 1 GPU): GPU tests, sanitizer passes on the fused RDB tail, ncu of K3/K4/K5, bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r02a_gpu.txt
( time timeout 2400 python -m pytest tests -m gpu -q -s -p no:cacheprovider ) > gpurun_out/r02a_pytest.log 2>&1
tail -n 25 gpurun_out/r02a_pytest.log
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_tail.py > gpurun_out/r02a_sanitize_${tool}.txt 2>&1
  tail -n 4 gpurun_out/r02a_sanitize_${tool}.txt
done
export BIN_B200_GRAPH=0
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:'pack_frames_kernel|convlstm_kernel|conv_igemm_kernel<16' -s 6 -c 8 -f -o gpurun_out/r02a_prof_k345 \
    python tools/run_window.py 2 > gpurun_out/r02a_ncu_k345.log 2>&1
tail -n 3 gpurun_out/r02a_ncu_k345.log
unset BIN_B200_GRAPH
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
tail -c 3000 gpurun_out/r02a_bench.json; tail -n 5 gpurun_out/r02a_bench.err
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/r02a_bench_ref.json 2> gpurun_out/r02a_bench_ref.err
cat gpurun_out/r02a_bench_ref.json; tail -n 5 gpurun_out/r02a_bench_ref.err
# CTA-pair fused tail: numerics on ragged shapes first (watchdog traps instead of hanging), then the A/B timing
BIN_B200_PAIR=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "rdb_tail_bit or rdb_shapes or cta_pair" -p no:cacheprovider > gpurun_out/r02a_pair_pytest.log 2>&1
tail -n 6 gpurun_out/r02a_pair_pytest.log
timeout 600 python tools/ab_pair.py > gpurun_out/r02a_ab_pair.txt 2>&1
cat gpurun_out/r02a_ab_pair.txt
ls -la gpurun_out | tail -15
