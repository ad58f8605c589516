python -m pytest tests/ -x -q -m gpu 2>&1 | tail -8
for wl in loftup train eval jbu; do python bench.py --workload $wl --steps 5 --warmup 3 2>/dev/null > gpurun_out/bench_${wl}_r1b.json; python -c "
import sys,json; d=json.loads(open('gpurun_out/bench_${wl}_r1b.json').read()); print('$wl', d['value'], d['unit'], 'e2e', d['e2e']['value'], 'ms', d['ms_per_step'], (d.get('roofline') or {}).get('frac'), d['clocks'])"; done
