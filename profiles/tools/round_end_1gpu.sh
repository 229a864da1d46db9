set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_r1z.log 2>&1; echo pytest rc=$?
for w in cfg3 cfg2 cfg4 cfg5 cfg1; do timeout 280 python bench.py --workload $w > gpurun_out/bench_r1z_$w.json 2> gpurun_out/bench_r1z_$w.err; echo $w rc=$?; done
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1z_reference_cfg3.json 2>/dev/null; echo ref rc=$?
timeout 200 python bench.py --impl reference --workload cfg1 --steps 3 --warmup 1 > gpurun_out/bench_r1z_reference_cfg1.json 2>/dev/null; echo ref1 rc=$?
timeout 200 python profiles/stage_times.py cfg3 cfg2 cfg4 cfg5 > gpurun_out/stage_times_r1z.log 2>&1
timeout 200 python profiles/run_reference_binary.py > gpurun_out/reference_binary_r1z.log 2>&1
for w in cfg3 cfg2; do timeout 300 ncu --set full --clock-control none --import-source on --kernel-name regex:"range_fft_kernel|doppler_fft|cfar|list_kernel|measure_kernel" --launch-skip 10 --launch-count 5 -f -o gpurun_out/ncu_r1z_stages_$w python profiles/prof_run.py $w > gpurun_out/ncu_r1z_stages_$w.log 2>&1; echo ncu $w rc=$?; done
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name regex:"range_fft_kernel|doppler_fft|cfar|list_kernel|measure_kernel|power_sum|merge" -c 400 --csv --log-file gpurun_out/launches_r1z_bench_cfg3.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-other > gpurun_out/launches_r1z_bench.log 2>&1; echo launches rc=$?
tail -3 gpurun_out/pytest_gpu_r1z.log
