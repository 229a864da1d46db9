# last check of the tree as committed: GPU tests, smoke, the default bench line
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_r2_last.log 2>&1; echo pytest rc=$?; tail -2 gpurun_out/pytest_gpu_r2_last.log
timeout 100 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 400 python bench.py > gpurun_out/bench_r2_last_cfg3.json 2> gpurun_out/bench_r2_last_cfg3.err; echo bench rc=$?
python -c "
import json
d=json.loads(open('gpurun_out/bench_r2_last_cfg3.json').read().strip().splitlines()[-1]); o=d['other_workloads']
print(round(d['value']), d['ms_per_step'], round(d['roofline']['frac'],3), round(d['e2e']['value']), d['clocks'], round(o['cfg2']['value']), round(o['cfg4']['value']), round(o['cfg5']['value']))"
