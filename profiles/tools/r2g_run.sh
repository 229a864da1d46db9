# cfg4 Doppler kernel with four warps per 8-row tile (every slot busy in both passes): MMW_K2_VARIANT=22 (3 stages, 2 CTAs/SM), 23 (2 stages, 3 CTAs/SM)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python profiles/sweep_env.py cfg4 MMW_K2_VARIANT=0,22,23,0,22,23 > gpurun_out/sweep_k2_512_r2g.log 2>&1; echo s512 rc=$?
cat gpurun_out/sweep_k2_512_r2g.log
for v in 22 23; do MMW_K2_VARIANT=$v timeout 600 python -m pytest tests/test_gpu_bench_paths.py -m gpu -q -x -k "cfg4" > gpurun_out/pytest_gpu_r2g_v$v.log 2>&1; echo pytest $v rc=$?; tail -2 gpurun_out/pytest_gpu_r2g_v$v.log; done
