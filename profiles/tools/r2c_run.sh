cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_r2c.log 2>&1; echo pytest rc=$?
tail -5 gpurun_out/pytest_gpu_r2c.log
bash profiles/tools/ab_lib.sh r2c cfg3 cfg2 cfg4 cfg5
timeout 400 python bench.py > gpurun_out/bench_r2c_cfg3.json 2> gpurun_out/bench_r2c_cfg3.err; echo bench rc=$?
