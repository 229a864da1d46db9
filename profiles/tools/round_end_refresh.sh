cd $GRAFT_REPO_ROOT
for w in cfg2 cfg4 cfg5; do timeout 280 python bench.py --workload $w > gpurun_out/bench_r1w_$w.json 2> gpurun_out/bench_r1w_$w.err; echo $w rc=$?; done
timeout 200 python profiles/stage_times.py cfg3 cfg2 cfg4 cfg5 > gpurun_out/stage_times_r1w.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name regex:"range_fft_kernel|doppler_fft|cfar|list_kernel|measure_kernel" --launch-skip 10 --launch-count 5 -f -o gpurun_out/ncu_r1w_stages_cfg2 python profiles/prof_run.py cfg2 > gpurun_out/ncu_r1w_stages_cfg2.log 2>&1; echo ncu rc=$?
