# Round-2 final 8-GPU run: the frame-sharded bench (cfg3 + cfg2/cfg4/cfg5 compact lines) and the C++ group example over all eight GPUs
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 > gpurun_out/bench_r2z_cfg3_n8.json 2> gpurun_out/bench_r2z_cfg3_n8.err; echo bench rc=$?
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -k "group_api" > gpurun_out/pytest_gpu_multi_r2z_n8.log 2>&1; echo pytest rc=$?
tail -2 gpurun_out/pytest_gpu_multi_r2z_n8.log
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_r2z_cfg3_n8.json').read().strip().splitlines()[-1])
print('cfg3 n8', round(d['value']), d['ms_per_step'], [round(x,4) for x in d['config']['ms_per_step_by_rank']], round(d['e2e']['value']))
for k,v in d['other_workloads'].items(): print(k, round(v['value']), v.get('ms_per_step'), round(v['e2e']['value']))
P
