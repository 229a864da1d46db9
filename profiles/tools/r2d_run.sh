# K2 frame order (MMW_K2_ORDER) and the fused front (MMW_FRONT=2) with the round-2c build; parity of the switched forms
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for i in 1 2; do timeout 200 python profiles/sweep_env.py cfg3 cfg2 cfg5 MMW_K2_ORDER=0,1; done > gpurun_out/sweep_k2_order_r2d.log 2>&1; echo order rc=$?
timeout 200 python profiles/sweep_env.py cfg3 MMW_FRONT=1,2 > gpurun_out/sweep_front_r2d.log 2>&1; echo front rc=$?
MMW_K2_ORDER=1 timeout 300 python bench.py --no-cpu-baseline --no-other > gpurun_out/bench_r2d_order1.json 2> gpurun_out/bench_r2d_order1.err; echo bench1 rc=$?
MMW_K2_ORDER=0 timeout 300 python bench.py --no-cpu-baseline --no-other > gpurun_out/bench_r2d_order0.json 2> gpurun_out/bench_r2d_order0.err; echo bench0 rc=$?
MMW_K2_ORDER=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_paths.py tests/test_gpu_properties.py -m gpu -q -x > gpurun_out/pytest_gpu_r2d_order1.log 2>&1; echo pytest rc=$?
cat gpurun_out/sweep_k2_order_r2d.log gpurun_out/sweep_front_r2d.log; tail -3 gpurun_out/pytest_gpu_r2d_order1.log
python - <<'P'
import json
for t in ("order1","order0"):
    d=json.loads(open(f"gpurun_out/bench_r2d_{t}.json").read().strip().splitlines()[-1])
    print(t, d["value"], d["ms_per_step"], d["config"]["ms_per_step_one_in_flight"], d["roofline"]["stage_ms"])
P
