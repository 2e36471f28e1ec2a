#!/bin/bash
# compute-sanitizer over the kernels of the later round-2 sessions: k_single_f64 (cooperative, beams 5 and 50),
# k_tc_vocab_dense, the dynamic candidate scores formed by k_score_nodes<DYN>, launch chaining (pdl_enter in every kernel)
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
echo "== memcheck"
timeout 1500 $S --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider --timeout 1400 \
  -k "(single_sentence_kernel and (small_tied- or small_dsoftmax_star or beam50)) or (tc_small and (dyn_top or tied_vs_top)) or (exact_dynamic and dyn_top)" > gpurun_out/sanitize_r2b_memcheck.log 2>&1; echo "rc=$?"
grep -E "passed|failed|ERROR SUMMARY|Invalid|Error" gpurun_out/sanitize_r2b_memcheck.log | tail -8
echo "== racecheck (shared-memory phases of k_single_f64)"
timeout 1500 $S --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider --timeout 1400 \
  -k "single_sentence_kernel and (small_tied- or beam50)" > gpurun_out/sanitize_r2b_racecheck.log 2>&1; echo "rc=$?"
grep -E "passed|failed|RACECHECK SUMMARY|hazard|Error" gpurun_out/sanitize_r2b_racecheck.log | tail -8
