cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 800 python scripts/explore_big_models.py 2>&1 | tail -14
