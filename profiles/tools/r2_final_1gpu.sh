# Round-2 final single-GPU evidence: tests, parity maxima, bench lines (default + cfg1 + the CPU arm), stage times, the ncu launch
# list of the bench command and `ncu --set full` stage captures of cfg2 / cfg4 (cfg3's is profiles/ncu_r2_stages_cfg3.*)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=r2z
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_$T.log 2>&1; echo pytest rc=$?
timeout 600 python -m pytest tests/test_gpu_bench_paths.py -m gpu -q -s > gpurun_out/parity_maxima_$T.log 2>&1; echo parity rc=$?
timeout 100 python __graft_entry__.py smoke > gpurun_out/smoke_$T.log 2>&1; echo smoke rc=$?
timeout 400 python bench.py > gpurun_out/bench_${T}_cfg3.json 2> gpurun_out/bench_${T}_cfg3.err; echo bench rc=$?
timeout 300 python bench.py --workload cfg1 > gpurun_out/bench_${T}_cfg1.json 2> gpurun_out/bench_${T}_cfg1.err; echo cfg1 rc=$?
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${T}_reference_cfg3.json 2>/dev/null; echo ref rc=$?
timeout 200 python profiles/stage_times.py cfg3 cfg2 cfg4 cfg5 > gpurun_out/stage_times_$T.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name regex:"range_fft_kernel|doppler_fft|cfar|list_kernel|measure_kernel|power_sum|merge" -c 400 --csv --log-file gpurun_out/launches_${T}_bench_cfg3.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-other > gpurun_out/launches_${T}_bench.log 2>&1; echo launches rc=$?
for w in cfg2 cfg4; do timeout 300 ncu --set full --clock-control none --import-source on --kernel-name regex:"range_fft_kernel|doppler_fft|doppler_extract|cfar|list_kernel|measure|rows_kernel|angle_fft" --launch-skip 9 --launch-count 9 -f -o gpurun_out/ncu_${T}_stages_$w python profiles/prof_run.py $w > gpurun_out/ncu_${T}_stages_$w.log 2>&1; echo ncu $w rc=$?; done
tail -3 gpurun_out/pytest_gpu_$T.log
