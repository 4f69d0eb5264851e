( timeout 1200 compute-sanitizer --tool racecheck --racecheck-report all python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "conv_slab_matches_fp64 and same3 and 3-16-16" ) > /tmp/race.txt 2>&1
grep -n "Race reported" -A12 /tmp/race.txt | head -60
