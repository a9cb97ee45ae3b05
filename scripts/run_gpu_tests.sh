make -C oracle >/dev/null 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo pytest_exit=$?
tail -30 gpurun_out/pytest_gpu.log
