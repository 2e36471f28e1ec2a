#!/bin/bash
# compute-sanitizer evidence for the kernels changed late in the round (CTA-pair GEMMs, staged LSTM epilogue,
# parallel prune, ScoreItem scoring, streaming submit/collect) + TC-vs-float64 parity at full size.
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
echo "== memcheck (exact + beam kernels)"
timeout 900 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_charrnn.py -m gpu -q -x -p no:cacheprovider \
  -k "small_tied or small_dsoftmax_star_vs or ragged or empty or tied_keep or charrnn_small" 2>&1 | tail -4
echo "== memcheck (tcgen05 / TMA / cluster kernels, streaming)"
timeout 900 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider \
  -k "tc_small or tc_gemm or stream" 2>&1 | tail -4
echo "== racecheck (prune / score / staged epilogue shared memory)"
timeout 900 $CS --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider \
  -k "tied_keep or (tc_small and small_tied-)" 2>&1 | tail -6
echo "== parity at scale"
timeout 600 python scripts/parity_at_scale.py 1024 2>&1 | tail -2
