mkdir -p gpurun_out/r2t
timeout 600 python -m pytest tests/test_gpu_dbrrt.py tests/test_gpu_tdbastar.py tests/test_gpu_backward.py -m gpu -x -q -s > gpurun_out/r2t/pytest_dbrrt.log 2>&1; echo "exit $?"; tail -12 gpurun_out/r2t/pytest_dbrrt.log
