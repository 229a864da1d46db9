# K1 tile heights re-measured now that tiles come from a counter (neighbouring chirp tiles of a slab in flight together)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{ timeout 200 python profiles/sweep_env.py cfg4 MMW_K1_VARIANT=0,6,0,6
timeout 200 python profiles/sweep_env.py cfg3 MMW_K1_VARIANT=0,5,7,0,5,7
timeout 200 python profiles/sweep_env.py cfg2 cfg5 MMW_K1_VARIANT=0,6,0,6; } > gpurun_out/sweep_k1_tiles_r2k.log 2>&1
cat gpurun_out/sweep_k1_tiles_r2k.log
