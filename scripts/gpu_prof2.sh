set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"stem_kernel|fuse_dw_kernel|pointwise_tc|depthwise_kernel" -o gpurun_out/prof_small -f python scripts/profile_small.py > gpurun_out/ncu_small.log 2>&1; tail -1 gpurun_out/ncu_small.log
du -sh gpurun_out
