set -x
python bench.py > gpurun_out/final_jbu.json 2> gpurun_out/final_jbu.err
python bench.py --workload loftup --steps 5 > gpurun_out/final_loftup.json 2> gpurun_out/final_loftup.err
python bench.py --workload train --steps 3 > gpurun_out/final_train.json 2> gpurun_out/final_train.err
python bench.py --workload eval --steps 3 > gpurun_out/final_eval.json 2> gpurun_out/final_eval.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/final_ref_jbu.json 2> gpurun_out/final_ref_jbu.err
# launch lists of the same commands (shares must agree with the event timings; absolute times are cold-cache, serialised)
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_jbu_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_jbu_final.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_loftup_final.csv python bench.py --workload loftup --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_loftup_final.log 2>&1
WHICH=attn python tools/prof_loftup_kernels.py && ncu --set full --clock-control none --import-source on -k regex:attention -c 1 -o gpurun_out/attn_v4 -f env WHICH=attn python tools/prof_loftup_kernels.py > gpurun_out/ncu_attn.log 2>&1
tail -c 300 gpurun_out/final_jbu.json; tail -2 gpurun_out/final_jbu.err; tail -c 200 gpurun_out/final_ref_jbu.json
