# ncu --set full (with source) of the five stage kernels of one cfg3 batch, round-2 build
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name regex:"range_fft_kernel|doppler_fft|cfar|list_kernel|measure_kernel" --launch-skip 10 --launch-count 5 -f -o gpurun_out/ncu_r2c_stages_cfg3 python profiles/prof_run.py cfg3 > gpurun_out/ncu_r2c_stages_cfg3.log 2>&1; echo ncu rc=$?
tail -3 gpurun_out/ncu_r2c_stages_cfg3.log
ls -la gpurun_out
