python -m pytest tests/test_gpu_tc.py tests/test_gpu_head_bwd.py -x -q -m gpu 2>&1 | tail -4
WHICH=conv python tools/bench_gemm_shapes.py 2>&1 | tail -4
