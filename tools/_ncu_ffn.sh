WHICH=fused ncu --set full --import-source on --clock-control none -k regex:ffn_fused_kernel --launch-skip 2 --launch-count 1 -o gpurun_out/r02_ffn_fused -f python tools/prof_ffn_fused.py > /dev/null 2>&1
ls -la gpurun_out/r02_ffn_fused.ncu-rep
