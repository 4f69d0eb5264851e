( timeout 900 python -m pytest tests/test_gpu_head.py -q -m gpu --timeout 300 -p no:cacheprovider -k "image_branch" -s ) 2>&1 | grep -E "^E  |passed|failed|^FAILED|image branch" | head -30
