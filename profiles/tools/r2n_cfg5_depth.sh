# cfg5 (one 256 x 128 x 12 frame per call): end-to-end calls/s against the number of calls in flight (bench.py --depth)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for d in 4 6 8 12; do
  timeout 120 python bench.py --workload cfg5 --depth $d --no-cpu-baseline > gpurun_out/bench_r2n_cfg5_depth$d.json 2>/dev/null
  python - $d <<'P'
import sys, json
d = json.loads(open(f"gpurun_out/bench_r2n_cfg5_depth{sys.argv[1]}.json").read().strip().splitlines()[-1])
print("depth", sys.argv[1], "device-resident", round(d["value"]), "e2e", round(d["e2e"]["value"]), "p50/p99 us", round(d["config"]["latency_us"]["p50"], 1), round(d["config"]["latency_us"]["p99"], 1))
P
done
