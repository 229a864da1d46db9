# the selective Doppler re-FFT + angle-FFT detection path (K4x) for NARROW arrays too (MMW_K4_VARIANT=2) against the per-detection kernel
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python profiles/sweep_env.py cfg2 cfg3 cfg5 MMW_K4_VARIANT=0,2,0,2 > gpurun_out/sweep_k4x_narrow_r2h.log 2>&1; echo sweep rc=$?
cat gpurun_out/sweep_k4x_narrow_r2h.log
MMW_K4_VARIANT=2 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_paths.py tests/test_gpu_properties.py -m gpu -q -x > gpurun_out/pytest_gpu_r2h_k4v2.log 2>&1; echo pytest rc=$?
tail -5 gpurun_out/pytest_gpu_r2h_k4v2.log
