python -m pytest tests/test_gpu_regions.py tests/test_gpu_tolerance.py tests/test_gpu_f64.py tests/test_gpu_cv.py tests/test_gpu_examples.py tests/test_gpu_full_size.py -x -q -m gpu 2>&1 | tail -3
BATCH=0 python profiles/run_full.py c4 2>&1 | tail -1
BATCH=0 python profiles/run_full.py c3 2>&1 | tail -1
python profiles/run_full.py tol 3e-12 2>&1 | tail -1
