# Round-end measurements: tests, bench (+cpu baseline), reference arm, ncu launch list (+DRAM
# bytes), ncu --set full of the representative kernels.  Outputs under gpurun_out/final_*.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
timeout 1200 python -m pytest tests -q -m gpu --timeout 300 2>&1 | tail -3 | tee gpurun_out/final_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/final_ops.json > gpurun_out/final_bench.log 2>&1; tail -1 gpurun_out/final_bench.log | cut -c1-300
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_ref.log 2>&1; tail -1 gpurun_out/final_ref.log | cut -c1-600
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1500 -c 700 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/final_ncu_bench.log 2>&1
tail -c 200 gpurun_out/final_ncu_bench.log
timeout 900 ncu --set full --clock-control none -k regex:"stem_kernel|fuse_dw_kernel|pointwise_tc|depthwise_kernel|sepconv|mbconv_front" -o gpurun_out/final_full -f python scripts/profile_kernels.py 1 > gpurun_out/final_ncu_full.log 2>&1; tail -1 gpurun_out/final_ncu_full.log
ncu -i gpurun_out/final_full.ncu-rep --page raw --csv > gpurun_out/final_full_raw.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep.tmp; [ $(du -sm gpurun_out | cut -f1) -gt 55 ] && rm -f gpurun_out/final_full.ncu-rep
du -sh gpurun_out; ls -la gpurun_out | tail -12
