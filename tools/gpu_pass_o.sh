#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_knn.py -m gpu -q -x 2>&1 | tail -2
timeout 300 python tools/microbench_knn.py 2>&1 | grep -E "tcgen05|redo" | head -8
