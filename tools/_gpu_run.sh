python -m pytest tests/test_gpu_models.py -x -q -m gpu 2>&1 | tail -4
python bench.py --workload eval --steps 3 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['unit'], d['ms_per_step'], d['gpu_launches'], d['noc'])"
