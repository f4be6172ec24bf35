set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pointwise_tc -c 2 -o gpurun_out/prof_pw -f python scripts/profile_kernels.py 1 > gpurun_out/ncu_pw.log 2>&1; tail -1 gpurun_out/ncu_pw.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:depthwise_kernel -c 3 -o gpurun_out/prof_dw -f python scripts/profile_kernels.py 1 > gpurun_out/ncu_dw.log 2>&1; tail -1 gpurun_out/ncu_dw.log
du -sh gpurun_out
