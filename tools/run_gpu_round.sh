timeout 600 python -m pytest tests -m gpu -x -q -k "ck or relabel" > gpurun_out/t68.log 2>&1; echo "pytest_exit=$?"; tail -n 12 gpurun_out/t68.log
