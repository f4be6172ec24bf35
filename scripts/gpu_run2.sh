set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
timeout 1200 python -m pytest tests -q -m gpu --timeout 300 2>&1 | tail -30 > gpurun_out/t_all.log; tail -12 gpurun_out/t_all.log
timeout 600 python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/ops_r1b.json > gpurun_out/bench2.log 2>&1; tail -3 gpurun_out/bench2.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log
timeout 900 ncu --set full --clock-control none --import-source on -o gpurun_out/prof_r1 -f python scripts/profile_kernels.py 1 > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
