python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
python bench.py --workload loftup --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['roofline'], d['clocks'])"
python bench.py --workload jbu --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['roofline'], d['clocks'])"
