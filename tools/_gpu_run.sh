python tools/prof_wgrad.py; B=8 python tools/prof_wgrad.py
python -m pytest tests/test_gpu_head_bwd.py -x -q -m gpu 2>&1 | tail -3
