cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
H2D_PROBE_TOPO=1 timeout 200 $TR --nproc-per-node 8 --master-port 29551 profiles/tools/h2d_probe_multi.py > gpurun_out/h2d_probe_n8.log 2>&1; echo rc=$?
timeout 200 $TR --nproc-per-node 4 --master-port 29552 profiles/tools/h2d_probe_multi.py > gpurun_out/h2d_probe_n4.log 2>&1; echo rc=$?
CUDA_VISIBLE_DEVICES=0,2,4,6 timeout 200 $TR --nproc-per-node 4 --master-port 29553 profiles/tools/h2d_probe_multi.py > gpurun_out/h2d_probe_n4_even.log 2>&1; echo rc=$?
grep probe gpurun_out/h2d_probe_n*.log
T0=$(date +%s)
timeout 600 $TR --nproc-per-node 8 --master-port 29554 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_r2_cfg3_n8.json 2> gpurun_out/bench_r2_cfg3_n8.err; echo rc=$? wall=$(( $(date +%s) - T0 ))s
tail -c 1500 gpurun_out/bench_r2_cfg3_n8.json
