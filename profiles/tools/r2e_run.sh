# K1 -> K2 hand-over as it happens in the chain: ncu WITHOUT its cache flush between kernels (--cache-control none), single-pass
# metrics only (time, DRAM bytes, L2 hit rate), K2 walking the batch forwards (MMW_K2_ORDER=0) and backwards (=1)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for o in 0 1; do
MMW_K2_ORDER=$o timeout 300 ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --kernel-name regex:"range_fft_kernel|doppler_fft|cfar|list_kernel|measure_kernel" --csv --log-file gpurun_out/launches_r2e_nocachectl_order$o.csv python profiles/prof_run.py cfg3 > gpurun_out/launches_r2e_order$o.log 2>&1; echo ncu $o rc=$?
done
grep -E "range_fft|doppler_fft" gpurun_out/launches_r2e_nocachectl_order0.csv | cut -d, -f1,5,11- | tail -24
echo; grep -E "range_fft|doppler_fft" gpurun_out/launches_r2e_nocachectl_order1.csv | cut -d, -f1,5,11- | tail -24
