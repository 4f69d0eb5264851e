python scripts/graph_vs_eager.py 256 2>&1 | grep -v Warning | tail -2
python scripts/graph_vs_eager.py 128 2>&1 | grep -v Warning | tail -1
python scripts/graph_vs_eager.py 32 2>&1 | grep -v Warning | tail -1
