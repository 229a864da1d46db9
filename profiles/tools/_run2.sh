cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run() { name=$1; shift; port=$((29580 + RANDOM % 300)); timeout 300 $TR --master-port $port bench.py --gpus 2 --steps 40 --warmup 3 --no-cpu-baseline --no-other "$@" > gpurun_out/bench_n2_x_$name.json 2> gpurun_out/bench_n2_x_$name.err; echo $name rc=$?; }
run peer --exchange peer
run nccl --exchange nccl
run peer2 --exchange peer
python - <<'PY'
import json
for f in ("peer","nccl","peer2"):
    try:
        d=json.loads(open(f"gpurun_out/bench_n2_x_{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"]), d["ms_per_step"], d["config"]["ms_per_step_by_rank"], d["config"]["host_issue_ms_per_step_by_rank"], d["config"]["detections_per_step"])
    except Exception as e: print(f, "failed", e, open(f"gpurun_out/bench_n2_x_{f}.err").read()[-1500:])
PY
