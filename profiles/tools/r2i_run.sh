# bench.py (cfg3 default + cfg2/cfg4/cfg5 compact lines) with the K4x detection path for narrow arrays (MMW_K4_VARIANT=2) against the default
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in 2 0 2 0; do MMW_K4_VARIANT=$v timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_r2i_k4v${v}_$RANDOM.json 2>/dev/null; echo bench $v rc=$?; done
python - <<'P'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_r2i_k4v*.json")):
    d=json.loads(open(f).read().strip().splitlines()[-1]); o=d["other_workloads"]
    print(f.split("/")[-1], "cfg3", round(d["value"]), round(d["e2e"]["value"]), "| cfg2", round(o["cfg2"]["value"]), round(o["cfg2"]["e2e"]["value"]), o["cfg2"]["stage_ms"]["list_kernel+measure_kernel"],
          "| cfg5", round(o["cfg5"]["value"]), round(o["cfg5"]["e2e"]["value"]), o["cfg5"]["latency_us"]["p50"], o["cfg5"]["latency_us"]["p99"], "| cfg4", round(o["cfg4"]["value"]))
P
