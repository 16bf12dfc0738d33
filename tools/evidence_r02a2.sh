#!/bin/bash
# Round-2 evidence bundle A2: locate the CTA-pair faults with memcheck, re-run the GPU tests after the ConvLSTM fix.
mkdir -p gpurun_out
for c in tail1 tailN conv3 convN; do
  BIN_B200_PAIR=1 timeout 300 compute-sanitizer --tool memcheck --print-limit 6 python tools/sanitize_pair.py $c > gpurun_out/r02a2_memcheck_$c.txt 2>&1
  echo "== $c"; grep -v "^=========     Host Frame\|^=========         in \|^=========     Saved host" gpurun_out/r02a2_memcheck_$c.txt | head -n 30
done
for c in tail1 tailN conv3 convN; do BIN_B200_PAIR=1 timeout 120 python tools/sanitize_pair.py $c 2>&1 | tail -n 2; done > gpurun_out/r02a2_pair_plain.txt 2>&1
cat gpurun_out/r02a2_pair_plain.txt
( time timeout 2400 python -m pytest tests -m gpu -q -s -p no:cacheprovider ) > gpurun_out/r02a2_pytest.log 2>&1
tail -n 25 gpurun_out/r02a2_pytest.log
export BIN_B200_GRAPH=0
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:'pack_frames_kernel|convlstm_kernel|conv_igemm_kernel<16' -s 4 -c 6 -f -o gpurun_out/r02a2_prof_k345 \
    python tools/run_window.py 2 > gpurun_out/r02a2_ncu_k345.log 2>&1
tail -n 3 gpurun_out/r02a2_ncu_k345.log
unset BIN_B200_GRAPH
ls -la gpurun_out | tail -12
