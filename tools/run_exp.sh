timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "narrow_bulk or project_fp32 or subsampled_hints" 2>&1 | tail -15
