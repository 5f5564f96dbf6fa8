timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu20.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu20.log
