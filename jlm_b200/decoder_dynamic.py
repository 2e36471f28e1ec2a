"""DynamicDecoder: incremental vocabulary selection (reference decoder/decoder_dynamic.py:18-194).

The reference grows a per-frame cumulative vocabulary, extends earlier frames' logits with the words
that appear later, re-normalises their softmax and re-scores every path from its head before each
pruning step.  All of that is a function of (a) each kept row's logits over the sentence's words and
(b) which words belong to lattice_vocab[i]; the device keeps a running log-sum-exp per row cut at the
frame boundaries and re-scores ancestors from those (jlm_beam.cu: k_dyn_prefix_lse, k_expand_prune).
"""
from . import _lib, lattice
from .decoder import Decoder


class DynamicDecoder(Decoder):
    dynamic = True

    def __init__(self, experiment_id=0, comp=0, device=0):
        super(DynamicDecoder, self).__init__(experiment_id=experiment_id, comp=comp, device=device)
        self.perf_log_fix_vocab = []                  # decoder_dynamic.py:27-28
        self.perf_log_fix_lattice_path_prob = []
        self._dyn = None

    def _build_lattice_vocab(self, frames, samples=0, top_sampling=False, random_sampling=False):
        # decoder_dynamic.py:30-46
        self._dyn = lattice.dynamic_vocab(frames, len(self.w2i), samples, top_sampling, random_sampling)
        self.lattice_vocab = None   # filled with the reference's end-of-decode state after the run

    def _decode_many(self, inputs, topN, beam_width, vocab_select, samples, top_sampling, random_sampling, backend,
                     timers):
        if not vocab_select:
            # the reference dereferences lattice_vocab=None here (decoder_dynamic.py:75)
            raise AttributeError("'NoneType' object has no attribute 'index' (DynamicDecoder needs vocab_select=True)")
        all_frames, dyn = [], []
        for text in inputs:
            frames = self._build_lattice(text, vocab_select=True, samples=samples, top_sampling=top_sampling,
                                         random_sampling=random_sampling)
            all_frames.append(frames)
            dyn.append(self._dyn)
        packed = lattice.PackedLattices(all_frames, dynamic=dyn)
        out = self._run(packed, _lib.DECODE_DYNAMIC, topN, beam_width, backend, timers=timers)
        self.lattice_vocab = lattice.dynamic_vocab_final(*dyn[-1])
        return all_frames, out

    def decode(self, input, topN=10, beam_width=10, vocab_select=False, samples=0, top_sampling=False,
               random_sampling=False, backend=_lib.BACKEND_AUTO):
        """decoder_dynamic.py:177-194"""
        all_frames, out = self._decode_many([input], topN, beam_width, vocab_select, samples, top_sampling,
                                            random_sampling, backend, timers=True)
        self.backward_lookup = lattice.to_backward_lookup(all_frames[0])
        info = self.last_info
        steps = max(int(info.n_steps) - 1, 1)          # the last frame is never stepped (T LM steps)
        self.perf_log_lstm += [info.ms_lstm * 1e-3 / steps] * steps
        self.perf_log_softmax += [info.ms_softmax * 1e-3 / steps] * steps
        # vocabulary fix-up and path re-scoring are fused into the softmax / beam kernels on the device
        self.perf_log_fix_vocab += [0.0] * steps
        self.perf_log_fix_lattice_path_prob += [info.ms_beam * 1e-3 / steps] * steps
        self.perf_sen += 1
        return out[0]

    def _array_mode(self, n_sent, vocab_select=True, samples=0, top_sampling=False, random_sampling=False):
        if not vocab_select:
            raise AttributeError("'NoneType' object has no attribute 'index' (DynamicDecoder needs vocab_select=True)")
        return _lib.DECODE_DYNAMIC, self._sample_ids(n_sent, samples, top_sampling, random_sampling)

    def decode_batch(self, inputs, topN=10, beam_width=10, vocab_select=True, samples=0, top_sampling=False,
                     random_sampling=False, backend=_lib.BACKEND_AUTO, native_lattice=True):
        inputs = list(inputs)
        if not inputs:
            return []
        if native_lattice and vocab_select:
            extra = self._sample_ids(len(inputs), samples, top_sampling, random_sampling)
            if not getattr(self, '_want_trace', False):
                out = self._run_texts(inputs, _lib.DECODE_DYNAMIC, extra, topN, beam_width, backend)
            else:
                packed = lattice.NativeLattices(self._native(), inputs, _lib.DECODE_DYNAMIC, extra)
                out = self._run(packed, _lib.DECODE_DYNAMIC, topN, beam_width, backend, timers=True)
        else:
            _, out = self._decode_many(inputs, topN, beam_width, vocab_select, samples, top_sampling,
                                       random_sampling, backend, timers=True)
        steps = self._log_batch_perf(len(out), last_frame_stepped=False)
        self.perf_log_fix_vocab += [0.0] * steps
        self.perf_log_fix_lattice_path_prob += [self.last_info.ms_beam * 1e-3 / steps] * steps
        return out
