set -x
ISP_TEST_REPORT=gpurun_out/r02_test_measurements.txt python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest_final.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02_pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke.log
python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "ref rc=$?"
B="python bench.py --workload loftup --steps 2 --warmup 1 --no-cpu-baseline --no-context"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/r02_launches_loftup_final.csv $B > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:attention_pair --launch-skip 18 --launch-count 1 -o gpurun_out/r02_ncu_attention_final -f $B > /dev/null 2>&1
ls -la gpurun_out/r02_ncu_attention_final.ncu-rep
