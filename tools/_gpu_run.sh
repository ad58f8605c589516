timeout 600 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
for c in 0 1; do for wl in loftup train; do ISP_GEMM_CTA2=$c timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cta2=$c $wl', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['clocks']['sm_mhz'])"; done; done
