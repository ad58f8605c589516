set -x
python -m pytest tests/test_gpu_tc.py tests/test_gpu_loftup.py -x -q -m gpu 2>&1 | tail -15
python bench.py --workload loftup --steps 5 --warmup 3 > gpurun_out/bench_loftup_ln.json 2> gpurun_out/bench_loftup_ln.err
tail -c 1500 gpurun_out/bench_loftup_ln.json; tail -5 gpurun_out/bench_loftup_ln.err
