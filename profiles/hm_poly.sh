for P in 0 4 8; do AS_HEADMEAN_POLY=$P timeout 200 python profiles/microbench.py 2>&1 | grep "headmean (" | sed "s/^/poly $P: /"; done
AS_HEADMEAN_POLY=8 timeout 300 python -m pytest tests/test_gpu_attention.py -x -q 2>&1 | tail -2
