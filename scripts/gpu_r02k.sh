python scripts/fullplay_bench.py 8192 2>&1 | tail -4
