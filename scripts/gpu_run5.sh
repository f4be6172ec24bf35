set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
timeout 1200 python -m pytest tests -q -m gpu --timeout 300 2>&1 | tail -30 > gpurun_out/t_all.log; tail -5 gpurun_out/t_all.log
timeout 300 python scripts/debug_nms.py 2>&1 | tail -5
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/ops_r1e.json > gpurun_out/bench5.log 2>&1; tail -1 gpurun_out/bench5.log | cut -c1-900
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nms_v5_fast -c 1 -o gpurun_out/prof_nms -f python scripts/debug_nms.py > gpurun_out/ncu_nms.log 2>&1; tail -2 gpurun_out/ncu_nms.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pointwise_tc|depthwise_kernel" -o gpurun_out/prof_kernels -f python scripts/profile_kernels.py 1 > gpurun_out/ncu_k.log 2>&1; tail -2 gpurun_out/ncu_k.log
ls -la gpurun_out; du -sh gpurun_out
