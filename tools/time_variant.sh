#!/bin/bash
# time the fused step (and the fused-gradient accuracy test) with the library given in $1 (default: in-tree build)
if [ -n "$1" ]; then export DSF_B200_LIB=$PWD/$1; fi
echo "lib=${DSF_B200_LIB:-in-tree}"
python tools/time_steps.py 4096 2>&1 | grep -E "chunks=[12] "
python -m pytest tests/test_gpu_bench_sizes.py -q -m gpu -k "fused_loss_gradient" -s 2>&1 | grep -E "passed|failed|err|rel" | head -12
