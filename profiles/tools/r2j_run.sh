cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_guard_bands.py tests/test_gpu_bench_paths.py -m gpu -q -x -s > gpurun_out/pytest_gpu_r2j.log 2>&1; echo pytest rc=$?
grep -E "passed|failed|cfg2|Error|error" gpurun_out/pytest_gpu_r2j.log | tail -12
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_r2j_cfg3.json 2>gpurun_out/bench_r2j_cfg3.err; echo bench rc=$?
python - <<'P'
import json
d=json.loads(open("gpurun_out/bench_r2j_cfg3.json").read().strip().splitlines()[-1]); o=d["other_workloads"]
print("cfg3", round(d["value"]), "cfg2", round(o["cfg2"]["value"]), o["cfg2"]["stage_ms"], o["cfg2"]["detect_path"], o["cfg2"]["hit_rows_retransformed"], o["cfg2"]["moved_frac_of_measured_hbm_peak"])
P
