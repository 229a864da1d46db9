# warp-private Doppler kernel at 512 points (cfg4) and at 128 points (cfg2/cfg5) after the round-2c staging change
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python profiles/sweep_env.py cfg4 MMW_K2_VARIANT=0,20,21,0,20,21 > gpurun_out/sweep_k2_512_r2f.log 2>&1; echo s512 rc=$?
timeout 300 python profiles/sweep_env.py cfg2 cfg5 MMW_K2_VARIANT=0,11,0,11 > gpurun_out/sweep_k2_128_r2f.log 2>&1; echo s128 rc=$?
cat gpurun_out/sweep_k2_512_r2f.log gpurun_out/sweep_k2_128_r2f.log
MMW_K2_VARIANT=20 timeout 600 python -m pytest tests/test_gpu_bench_paths.py tests/test_gpu_parity.py -m gpu -q -x -k "cfg4 or 512 or imaging or wide" > gpurun_out/pytest_gpu_r2f_v20.log 2>&1; echo pytest rc=$?
tail -4 gpurun_out/pytest_gpu_r2f_v20.log
