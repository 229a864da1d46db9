set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_r2a.log 2>&1; echo pytest rc=$?
timeout 400 python bench.py > gpurun_out/bench_r2a_cfg3.json 2> gpurun_out/bench_r2a_cfg3.err; echo bench rc=$?
timeout 200 python profiles/stage_times.py cfg3 cfg2 cfg4 cfg5 > gpurun_out/stage_times_r2a.log 2>&1
tail -3 gpurun_out/pytest_gpu_r2a.log
