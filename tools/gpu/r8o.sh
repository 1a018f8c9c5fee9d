set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --workload cfg4 --steps 10 --warmup 3 --no-cpu-baseline --no-sub --no-e2e --no-parity 2>/dev/null | tail -c 700
