set -x
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
bash tools/ncu_capture.sh r2z encode_fft_kernel xcorr_pair_kernel scan_score_kernel
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2z.csv python bench.py --pairs 65536 --target-total 4294967296 --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/launches_r2z.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pair_fused_kernel -s 4 -c 1 -f -o gpurun_out/prof_pair_fused_kernel_r2z python bench.py --pairs 32768 --target-total 4294967296 --steps 1 --warmup 3 --no-cpu-baseline --no-extras --fuse-pairs 1 > gpurun_out/ncu_fused_r2z.log 2>&1
tail -c 600 gpurun_out/bench_default.err; tail -c 300 gpurun_out/bench_reference.json
