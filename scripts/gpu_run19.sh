cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -k "lite3" 2>&1 | tail -25
