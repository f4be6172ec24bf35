set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 660 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -c 300 gpurun_out/ncu_bench.log
