#!/bin/bash
# compute-sanitizer over the kernels added in round 2 (through the C ABI via pytest)
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
echo "== memcheck: rs kernel (odd shapes, D-softmax*), KS stage-1, guard queue / paths / pool, keep-all, shared word rows"
timeout 2400 $S --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider --timeout 2000 \
  -k "odd_shapes and dsoftmax_star or near_tie or beam_width_none or (tc_small and (dyn_top or tied_vs or small_tied-)) or arrays_and_sharded or static_vocab_word" > gpurun_out/sanitize_r2_memcheck.log 2>&1; echo "rc=$?"
grep -E "passed|failed|ERROR SUMMARY|Invalid|Error" gpurun_out/sanitize_r2_memcheck.log | tail -8
echo "== racecheck: shared-memory kernels (k_prune_block guard section, k_dyn_prefix_lse scan, k_stream_f64, k_pool_lse_subset)"
timeout 2400 $S --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider --timeout 2000 \
  -k "near_tie or (exact_dynamic and dyn_top) or tied_keep" > gpurun_out/sanitize_r2_racecheck.log 2>&1; echo "rc=$?"
grep -E "passed|failed|RACECHECK SUMMARY|hazard|Error" gpurun_out/sanitize_r2_racecheck.log | tail -8
