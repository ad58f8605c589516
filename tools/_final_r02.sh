set -x
ISP_TEST_REPORT=gpurun_out/r02_test_measurements.txt python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest_final.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02_pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke.log
python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "ref rc=$?"
bash tools/_prof_r02.sh > gpurun_out/r02_prof.log 2>&1
J="python bench.py --workload jbu --steps 2 --warmup 1 --no-cpu-baseline --no-context"
ncu --set full --import-source on --clock-control none -k regex:bilinear_ac_march --launch-skip 1 --launch-count 1 -o gpurun_out/r02_ncu_resize_march -f $J > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:adaptive_conv_v3 --launch-skip 7 --launch-count 1 -o gpurun_out/r02_ncu_adaptive_conv_512 -f $J > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
