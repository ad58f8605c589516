python tools/prof_loftup_gemms.py
WHICH=out_proj ncu --set full --import-source on --clock-control none -k regex:gemm_tc_kernel --launch-skip 2 --launch-count 1 -o gpurun_out/r02_gemm_outproj -f python tools/prof_loftup_gemms.py > /dev/null 2>&1
WHICH=ff1 ncu --set full --import-source on --clock-control none -k regex:gemm_tc_kernel --launch-skip 2 --launch-count 1 -o gpurun_out/r02_gemm_ff1 -f python tools/prof_loftup_gemms.py > /dev/null 2>&1
ls -la gpurun_out/r02_gemm_*.ncu-rep
