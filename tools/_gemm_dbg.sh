timeout 300 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -3
for d in 0 8; do echo "dbg=$d"; ISP_GEMM_DBG=$d python tools/prof_loftup_gemms.py; done
echo "nowres"; ISP_GEMM_NO_WRES=1 python tools/prof_loftup_gemms.py
python tools/_chunk_test.py 2>&1 | tail -2
