python -m pytest tests/test_gpu_cv.py tests/test_gpu_fubini.py tests/test_gpu_examples.py tests/test_gpu_full_size.py -x -q -m gpu 2>&1 | tail -4
BATCH=0 python profiles/run_full.py c4 2>&1 | tail -2
BATCH=0 python profiles/run_full.py fub 2>&1 | tail -1
