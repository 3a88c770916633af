mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_dist.py -m gpu -q > gpurun_out/s3_pytest_dist_n2.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/s3_pytest_dist_n2.log
