set -x
B="python bench.py --workload loftup --steps 2 --warmup 1 --no-cpu-baseline --no-context"
$B > gpurun_out/r02_bench_loftup_short.json 2> gpurun_out/r02_bench_loftup_short.err || { tail -5 gpurun_out/r02_bench_loftup_short.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/r02_launches_loftup_final.csv $B > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:attention --launch-skip 18 --launch-count 1 -o gpurun_out/r02_ncu_attention_final -f $B > /dev/null 2>&1
WHICH=ff1 ncu --set full --import-source on --clock-control none -k regex:gemm_tc_kernel --launch-skip 2 --launch-count 1 -o gpurun_out/r02_ncu_gemm_ff1_final -f python tools/prof_loftup_gemms.py > /dev/null 2>&1
J="python bench.py --workload jbu --steps 2 --warmup 1 --no-cpu-baseline --no-context"
$J > gpurun_out/r02_bench_jbu_short.json 2>/dev/null && ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_jbu_final.csv $J > /dev/null 2>&1
ls -la gpurun_out/r02_launches_* gpurun_out/r02_ncu_*final*
