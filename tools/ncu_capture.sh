#!/bin/bash
# One `ncu --set full` capture per kernel of the path (run under gpurun; reports land in gpurun_out/).
#   tools/ncu_capture.sh <tag> [kernel-regex ...]
tag=${1:-r1}; shift
kernels=${@:-encode_fft_kernel xcorr_pair_kernel scan_score_kernel}
for k in $kernels; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f \
      -o gpurun_out/prof_${k}_${tag} python bench.py --pairs 32768 --target-total 4294967296 --steps 1 --warmup 3 --no-cpu-baseline --no-extras \
      > gpurun_out/ncu_${k}_${tag}.log 2>&1
  tail -2 gpurun_out/ncu_${k}_${tag}.log
done
