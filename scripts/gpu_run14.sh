set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 600 python scripts/bench_effnetv2.py efficientnetv2-s 128 384 > gpurun_out/effnetv2_s.log 2>&1; tail -1 gpurun_out/effnetv2_s.log
