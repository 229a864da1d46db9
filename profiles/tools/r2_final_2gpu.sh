# Round-2 final 2-GPU check: the multi-GPU tests (NCCL gather, peer exchange, group API from C++ and Python) and the frame-sharded bench
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/pytest_gpu_multi_r2z_n2.log 2>&1; echo pytest rc=$?
tail -3 gpurun_out/pytest_gpu_multi_r2z_n2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 > gpurun_out/bench_r2z_cfg3_n2.json 2> gpurun_out/bench_r2z_cfg3_n2.err; echo bench rc=$?
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2z_reference_n2.json 2>/dev/null; echo ref rc=$?
tail -c 600 gpurun_out/bench_r2z_cfg3_n2.json
