# Round-2 end measurements: smoke, whole GPU suite, bench (+cpu baseline) for the headline config,
# the reference arm, bench lines of the other BASELINE configs, segment timing, ncu launch list
# (+DRAM bytes) of the bench command, ncu --set full of the representative kernels.
# Outputs under gpurun_out/final_*; scripts/make_profiles_r2.py turns them into profiles/r2_*.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 2>&1 | tail -4 | tee gpurun_out/final_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/final_ops_d0.json > gpurun_out/final_bench_d0.log 2>&1; tail -1 gpurun_out/final_bench_d0.log | cut -c1-300
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/final_ref.log 2>&1; tail -1 gpurun_out/final_ref.log | cut -c1-400
for c in d4 d7x v2s; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 5 --profile-out gpurun_out/final_ops_$c.json > gpurun_out/final_bench_$c.log 2>&1
  tail -1 gpurun_out/final_bench_$c.log | cut -c1-260
done
timeout 300 python scripts/time_segments.py d0 gpurun_out/final_segments_d0.json 2>&1 | tail -3
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1500 -c 700 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/final_ncu_bench.log 2>&1
tail -c 200 gpurun_out/final_ncu_bench.log
timeout 900 ncu --set full --clock-control none -k regex:"stem_tc_kernel|fuse_dw_kernel|pointwise_tc|dw_tile_kernel|depthwise_kernel|sepconv|mbconv_front" -o gpurun_out/final_full -f python scripts/profile_kernels.py 1 > gpurun_out/final_ncu_full.log 2>&1; tail -1 gpurun_out/final_ncu_full.log
ncu -i gpurun_out/final_full.ncu-rep --page raw --csv > gpurun_out/final_full_raw.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep.tmp; [ $(du -sm gpurun_out | cut -f1) -gt 55 ] && rm -f gpurun_out/final_full.ncu-rep gpurun_out/d_full.ncu-rep
du -sh gpurun_out; ls -la gpurun_out | tail -12
