python -m pytest tests/test_gpu_jbu.py -x -q -m gpu 2>&1 | tail -2
python tools/bench_jbu_kernels.py 2>&1 | grep "bicubic"
