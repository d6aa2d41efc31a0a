python -m pytest tests -x -q -m gpu 2>&1 | tail -8
python tools/bench_configs.py --only cfg4 --short 2>&1 | tail -1 | cut -c1-200
