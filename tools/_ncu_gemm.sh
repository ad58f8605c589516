ISP_GEMM_DBG=8 WHICH=q_proj ncu --set full --import-source on --clock-control none -k regex:gemm_tc_kernel --launch-skip 2 --launch-count 1 -o gpurun_out/r02_gemm_qproj_dbg8 -f python tools/prof_loftup_gemms.py > /dev/null 2>&1
WHICH=q_proj ncu --set full --import-source on --clock-control none -k regex:gemm_tc_kernel --launch-skip 2 --launch-count 1 -o gpurun_out/r02_gemm_qproj_v2 -f python tools/prof_loftup_gemms.py > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
