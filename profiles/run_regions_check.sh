python -m pytest tests -x -q -m gpu 2>&1 | tail -4
BATCH=0 python profiles/run_full.py c4 2>&1 | tail -2
BATCH=0 python profiles/run_full.py c3 2>&1 | tail -2
python profiles/run_full.py tol 3e-12 2>&1 | tail -1
