#!/bin/bash
# compute-sanitizer on the kernels added after gpu_sanitize.sh: k_tc_vocab_logits (manually swizzled operand tiles),
# scan-based k_dyn_prefix_lse, k_tc_gather_rows / per-slot h split, the state pool.
CS=/usr/local/cuda/bin/compute-sanitizer
echo "== memcheck (vocabulary-selection modes on the tensor-core back end, per-slot h split)"
timeout 900 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider \
  -k "tc_small or (tc_full_size and cfg4) or odd_shapes" 2>&1 | tail -3
echo "== memcheck (state pool, char-RNN decoder)"
timeout 900 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_charrnn.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
echo "== racecheck (k_dyn_prefix_lse scan buffer, k_prune, k_tc_vocab_logits staging)"
timeout 900 $CS --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider \
  -k "tc_small and (dyn_top or tied_vs)" 2>&1 | tail -5
