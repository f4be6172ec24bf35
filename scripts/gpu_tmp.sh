cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 600 python -m pytest tests -q -m gpu --timeout 200 -x -k "mbconv or sepconv or network_parity or detect" 2>&1 | tail -4
AB_POSTPROCESS=1 timeout 300 python scripts/ab_engine.py with_nms= 2>&1 | tail -2
AB_POSTPROCESS=0 timeout 300 python scripts/ab_engine.py net_only= 2>&1 | tail -2
