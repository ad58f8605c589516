python -m pytest tests/test_gpu_jbu.py tests/test_gpu_models.py tests/test_gpu_attention.py -x -q -m gpu 2>&1 | tail -5
python tools/bench_jbu_kernels.py > gpurun_out/jbu_kernels.txt 2>&1; tail -30 gpurun_out/jbu_kernels.txt
WORKLOAD=jbu BATCH=16 python tools/stage_times.py 2>&1 | grep -v "n=   1     0.0" | tail -30
