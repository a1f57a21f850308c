set -x
python -m pytest tests/test_enc_gpu.py -x -q 2>&1 | tail -30
