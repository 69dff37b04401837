run() { for i in 1 2 3 4; do timeout 100 python -m pytest tests/test_gpu_gemm.py -q -k "4096-4096-512-3 or 4096-11008-16-3 or 1024-4096-130-3 or 11008-4096-32-3" 2>&1 | tail -1; done; }
echo BASE; run
echo NO_PDL; GGB200_NO_PDL=1 run
echo LATE_RELEASE; GGB200_GEMM_DBG=4 run
echo CG1; GGB200_GEMM_CG=1 run
