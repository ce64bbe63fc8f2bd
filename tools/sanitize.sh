#!/bin/bash
# compute-sanitizer over GPU parity tests (run under gpurun); summaries land in gpurun_out/sanitizer_<tool>_<tag>.log
tag=${1:-r2}
SEL='synthetic_edge or samples_stage or ragged or periodic or split_transform or guide_mode or pool_overflow or pruned_scan_equals_exhaustive_scan[20000.0] or random_pairs_against_oracle[4294967296.0] or sharded_handles'
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "$SEL" > gpurun_out/sanitizer_memcheck_${tag}.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/sanitizer_memcheck_${tag}.log
SEL2='synthetic_edge or periodic_sequences_overflow_unit_list[4096] or split_transform or random_pairs_against_oracle[4294967296.0]'
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "$SEL2" > gpurun_out/sanitizer_racecheck_${tag}.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/sanitizer_racecheck_${tag}.log
tail -n 4 gpurun_out/sanitizer_memcheck_${tag}.log; tail -n 4 gpurun_out/sanitizer_racecheck_${tag}.log
# the routes added in the second session of round 2 (three-channel form, fused kernel + fall-back lists, cluster kernel)
for tool in memcheck racecheck; do
  compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_paths.py > gpurun_out/sanitizer_${tool}_paths_${tag}.log 2>&1
  echo "$tool rc=$?" >> gpurun_out/sanitizer_${tool}_paths_${tag}.log
  tail -n 4 gpurun_out/sanitizer_${tool}_paths_${tag}.log
done
