# Round 2, call D: fused-front A/B with the new kernels, ncu --set full of the new kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 300 -x -k "pointwise" 2>&1 | tail -4
for ff in 1 0; do
  EDET_FUSE_FRONT=$ff timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/d_ops_d0_ff$ff.json > gpurun_out/d_bench_d0_ff$ff.log 2>&1
  echo "fuse_front=$ff: $(tail -1 gpurun_out/d_bench_d0_ff$ff.log | cut -c1-230)"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dw_tile_kernel|pointwise_tc_kernel|fuse_dw_kernel|sepconv_direct|stem_kernel" -o gpurun_out/d_full -f python scripts/profile_kernels.py 1 > gpurun_out/d_ncu_full.log 2>&1; tail -2 gpurun_out/d_ncu_full.log
ncu -i gpurun_out/d_full.ncu-rep --page raw --csv > gpurun_out/d_full_raw.csv 2>/dev/null
ls -la gpurun_out/d_full.ncu-rep
