for wl in jbu loftup; do python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/err_$wl.txt > gpurun_out/bench_${wl}_r1c.json; tail -3 gpurun_out/err_$wl.txt; python -c "
import sys,json; d=json.loads(open('gpurun_out/bench_${wl}_r1c.json').read()); print('$wl', d['value'], d['unit'], 'e2e', d['e2e'], 'ms', d['ms_per_step'], (d.get('roofline') or {}).get('frac'), d['clocks'])"; done
python -m pytest tests/test_gpu_models.py -x -q -m gpu 2>&1 | tail -3
