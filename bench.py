#!/usr/bin/env python
"""bench.py - kana chars/sec decoded (BASELINE.json metric).

Headline = configs[1] (cfg2): V=50000 H=512 embed=256, standard (tied) softmax, beam=10, 1 GPU; the other
BASELINE configs (cfg3 D-softmax*, cfg4 DynamicDecoder, cfg5 V=100k H=1024 beam 50) ride along as short runs in
the `workloads` array of the same JSON line, each with its own value / e2e / roofline / parity counts.

A "step" is one pass of the hot path over one batch of synthetic sentences: S sentences decoded in lock-step
per GPU.  Independent sentences are partitioned across ranks, no collective on the data path:
  weak scaling   (`value`, `e2e`)  : S sentences per GPU, private to the rank
  strong scaling (`strong`, cfg4/5): ONE fixed set of S sentences through jlm_b200.shard.decode_sharded
                                     (length-balanced partition, decode, packed NCCL all_gather of the n-best)

    python bench.py [--gpus N] [--steps K] [--warmup W] [--sentences S] [--impl ours|reference]
                    [--workload cfg2] [--extra cfg3,cfg4,cfg5|none] [--scaling weak|strong]

Prints ONE JSON line (rank 0).  Keys follow the driver contract:
  value      chars/s, lattices resident in HBM, CUDA-event time around run + n-best fetch (near-tie guard work
             included), max over ranks
  e2e        chars/s through jlm_decode_texts_submit/_collect with HOST buffers (H2D + D2H inside)
  roofline   the dominant kernel (output-projection GEMM + online-LSE epilogue): algorithmic FLOP/s over its
             CUDA-event duration vs the measured bf16 peak; `gate` holds the same for the LSTM gate GEMM
  guard      near-tie guard statistics: flagged_fraction, pairs re-scored, sentences re-decoded in float64
  cpu_baseline  the CPU oracle (numpy port of the reference) on a bounded sample, same host
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

if '--impl' in sys.argv and 'reference' in sys.argv:
    # the CPU arm may use every host thread; torchrun exports OMP_NUM_THREADS=1, which would pin
    # numpy's BLAS to one core - undo that before numpy loads its thread pool
    for _k in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS'):
        os.environ[_k] = str(os.cpu_count() or 1)

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

METRIC = 'kana chars/sec decoded at V=50k H=512 beam=10'
MIN_LEN = 20
# BASELINE.json configs[1..4]; configs[1] (cfg2) is the one the metric is quoted on and the default.
WORKLOADS = {
    'cfg2': dict(desc='cfg2: V=50000 H=512 E=256 tied standard softmax, beam=10, topN=10, >=20 kana/sentence',
                 V=50000, H=512, E=256, mode='tied', segments=None, beam=10, topn=10, dynamic=False),
    'cfg3': dict(desc='cfg3: V=50000 H=512 D-softmax* segs (200,0,12k)/(100,12k,30k)/(50,30k,50k), beam=10',
                 V=50000, H=512, E=256, mode='dsoftmax_star',
                 segments=[[200, 0, 12000], [100, 12000, 30000], [50, 30000, None]], beam=10, topn=10, dynamic=False),
    'cfg4': dict(desc='cfg4: V=50000 H=512 E=256 tied, DynamicDecoder (incremental vocab, samples=200 top), beam=20',
                 V=50000, H=512, E=256, mode='tied', segments=None, beam=20, topn=10, dynamic=True, samples=200),
    'cfg5': dict(desc='cfg5: V=100000 H=1024 D-softmax* segs (256,0,4k)/(128,4k,12k)/(64,12k,100k), beam=50',
                 V=100000, H=1024, E=256, mode='dsoftmax_star',
                 segments=[[256, 0, 4000], [128, 4000, 12000], [64, 12000, None]], beam=50, topn=10, dynamic=False),
}
CPU_NOTE = ('numpy port of the reference (oracle/jlm_oracle.py); measured equal to /root/reference on the cfg2 workload '
            '(18.7 vs 18.9 chars/s on 8 vCPU, same top-1) - the Python reference itself cannot travel to the GPU box')


def flops_per_row(wl):
    """Algorithmic flops of one LM row: (gate GEMM, full-vocabulary output GEMMs)."""
    V, H = wl['V'], wl['H']
    if wl['mode'] == 'dsoftmax_star':
        segs = [(sz, s, V if e is None else e) for sz, s, e in wl['segments']]
        e_in = segs[0][0]
        proj = sum(2.0 * sz * (e - s) for sz, s, e in segs)
    else:
        e_in = wl['E']
        proj = 2.0 * wl['E'] * V
    return 2.0 * (e_in + H) * 4 * H, proj


def peaks():
    p = os.path.join(REPO, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get('bf16_tflops_sustained', 1431.6), d.get('bf16_tflops', 1644.3), d.get('hbm_gbs', 6454.6), 'measured'
    return 1400.0, 1590.0, 6650.0, 'fallback'


def make_inputs(root, wl, n_sent, seed):
    from jlm_b200 import synth
    cfg, weights, lexicon, reading_dict = synth.make_experiment(root, 1, wl['V'], wl['H'], wl['E'], wl['mode'],
                                                                segments=wl['segments'], seed=0)
    sents = synth.make_sentences(lexicon, n_sent, min_len=MIN_LEN, seed=seed, vocab_size=wl['V'])
    return cfg, weights, lexicon, reading_dict, sents


def oracle_decoder(wl, cfg, weights, lexicon, reading_dict):
    """(prepare(text) -> state, run(state) -> n-best): the CPU oracle split so that only the
    lattice -> n-best region is timed (decode() minus _build_lattice)."""
    from oracle import jlm_oracle as O
    ora = O.OracleDecoder(cfg, weights, lexicon, reading_dict, dynamic=wl['dynamic'])

    def prepare(text):
        fr = O.build_lattice(text, ora.w2i, lexicon, reading_dict)
        if wl['dynamic']:
            return fr, O.dynamic_lattice_vocab(fr, len(ora.w2i), wl['samples'], True, False)
        return fr, None

    def run(state):
        fr, lv = state
        if wl['dynamic']:
            return O.decode_dynamic(ora.model, fr, {k: list(v) for k, v in lv.items()}, wl['topn'], wl['beam'])
        return O.decode_static(ora.model, fr, wl['topn'], wl['beam'], None)

    return prepare, run


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu = gpu
        self.rows = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append((time.perf_counter(), [x.strip() for x in line.split(',')]))
                if self.stop_flag:
                    break
        except Exception:
            pass

    def stop(self, windows):
        """Summary of the samples taken inside the (t_lo, t_hi) windows (the GPU is under load for all of them)."""
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        rows = [r for t, r in self.rows if any(lo <= t <= hi for lo, hi in windows)]
        sm = [float(r[0]) for r in rows if r and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in rows if len(r) > 1 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({names[i] for r in rows if len(r) >= 6 for i in range(4) if r[2 + i].lower() == 'active'})
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(sm)}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU algorithm (numpy oracle port - the Python reference
    itself cannot travel to the GPU box) on the host cores, bounded sample of the same workload.
    Every step decodes a different slice of a pool of >= 32 sentences."""
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    root = tempfile.mkdtemp(prefix='jlm_bench_ref_')
    per_step = max(1, args.ref_sentences)
    pool = max(32, per_step * min(args.steps, 8))
    cfg, weights, lexicon, reading_dict, sents = make_inputs(root, wl, pool, seed=100)
    prepare, run = oracle_decoder(wl, cfg, weights, lexicon, reading_dict)
    states = [prepare(s) for s in sents]

    def step(i):
        lo = (i * per_step) % pool
        idx = [(lo + k) % pool for k in range(per_step)]
        for j in idx:
            run(states[j])
        return sum(len(sents[j]) for j in idx), idx

    for i in range(args.warmup if args.warmup < 2 else 1):
        step(i)
    chars, seen = 0, set()
    t0 = time.perf_counter()
    for i in range(args.steps):
        c, idx = step(i)
        chars += c
        seen.update(idx)
    dt = time.perf_counter() - t0
    val = chars / dt
    sample = ('%d sentences per step, a different slice of a %d-sentence pool each step (%d distinct sentences, %d chars timed), '
              'lattice->n-best (decode minus _build_lattice)' % (per_step, pool, len(seen), chars))
    line = {'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': 'chars/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': wl['desc'], 'sentences_per_step': per_step, 'distinct_sentences_timed': len(seen)},
            'cpu_baseline': {'value': val, 'unit': 'chars/s', 'cores': os.cpu_count(), 'kind': 'port', 'sample': sample,
                             'note': CPU_NOTE},
            'e2e': {'value': val, 'unit': 'chars/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'note': CPU_NOTE, 'gpu_launches': 0}
    print(json.dumps(line), flush=True)


class Env(object):
    """Process-level state shared by the workloads of one bench invocation."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.rank = int(os.environ.get('RANK', '0'))
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.local = int(os.environ.get('LOCAL_RANK', '0'))
        if not torch.cuda.is_available():
            raise SystemExit('bench.py needs a CUDA device: libjlm_b200 has no CPU fallback')
        torch.cuda.set_device(self.local)
        if self.world > 1:
            os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
            dist.init_process_group('nccl', device_id=torch.device('cuda', self.local))
            # ranks share the host: lattice / plan threads are split between them (shard.host_threads_per_rank)
            from jlm_b200 import shard
            os.environ.setdefault('JLM_HOST_THREADS', str(min(8, shard.host_threads_per_rank(self.world))))
        from jlm_b200 import _lib
        self._lib = _lib
        self.lib = _lib.load()
        self.stream = torch.cuda.current_stream()
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')   # > 126 MB L2
        self.sampler = ClockSampler(self.local) if self.rank == 0 else None
        if self.sampler:
            self.sampler.start()
        self.windows = []

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, x, op):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device='cuda')
        self.dist.all_reduce(t, op=getattr(self.dist.ReduceOp, op))
        return float(t.item())


def bench_workload(env, name, steps, warmup, cpu_sentences, headline):
    """One workload on this rank's GPU: device-resident arm, roofline pass, e2e arm, (strong-scaling arm), CPU leg."""
    import jlm_b200
    from jlm_b200 import config, lattice, shard
    args, lib, _lib, torch = env.args, env.lib, env._lib, env.torch
    rank, world, local = env.rank, env.world, env.local
    wl = WORKLOADS[name]
    TOPN, BEAM, WORKLOAD = wl['topn'], wl['beam'], wl['desc']
    MODE = _lib.DECODE_DYNAMIC if wl['dynamic'] else _lib.DECODE_FULL
    root = tempfile.mkdtemp(prefix='jlm_bench_%s_r%d_' % (name, rank))
    cfg, weights, lexicon, reading_dict, sents = make_inputs(root, wl, args.sentences, seed=100 + rank)
    config.set_root(root)
    dec = (jlm_b200.DynamicDecoder if wl['dynamic'] else jlm_b200.Decoder)(1, device=local)
    extra = np.tile(np.arange(wl['samples'], dtype=np.int32), (len(sents), 1)) if wl['dynamic'] else None
    n_extra = wl['samples'] if wl['dynamic'] else 0
    hdl = dec.model._handle
    stream = env.stream
    _lib.check(lib.jlm_set_stream(hdl, C.c_void_p(stream.cuda_stream)))
    nlex = dec._native()
    t_lat = time.perf_counter()
    packed = lattice.NativeLattices(nlex, sents, MODE, extra)          # jlm_lattice_build (host C++)
    t_lat = time.perf_counter() - t_lat
    lb = packed.c_struct()
    chars = sum(len(s) for s in sents)
    S = packed.n_sent
    max_len = int(packed.sent_len.max()) + 1
    scores = np.empty((S, TOPN))
    n_paths = np.empty(S, dtype=np.int32)
    path_len = np.empty((S, TOPN), dtype=np.int32)
    path_nodes = np.zeros((S, TOPN, max_len), dtype=np.int32)
    nb = _lib.NBest()
    nb.top_n, nb.max_len = TOPN, max_len
    nb.scores, nb.n_paths = _lib.ptr(scores, C.c_double), _lib.ptr(n_paths, C.c_int32)
    nb.path_len, nb.path_nodes = _lib.ptr(path_len, C.c_int32), _lib.ptr(path_nodes, C.c_int32)

    # ---------------- device-resident arm: lattices uploaded once, K timed (run + fetch) ----------------
    batch = C.c_void_p()
    _lib.check(lib.jlm_batch_upload(hdl, C.byref(lb), BEAM, TOPN, MODE, args.backend, C.byref(batch)))
    if args.profile:
        for _ in range(1 + steps):
            _lib.check(lib.jlm_batch_run(batch))
        torch.cuda.synchronize()
        _lib.check(lib.jlm_batch_destroy(batch))
        return None
    info = _lib.BatchInfo()
    for _ in range(max(warmup, 3)):
        _lib.check(lib.jlm_batch_run(batch))
        _lib.check(lib.jlm_batch_fetch(batch, C.byref(nb)))
    env.barrier()
    t_load0 = time.perf_counter()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    env.barrier()
    wall0 = time.perf_counter()
    flagged = pairs = rerun = 0
    for a, b in ev:
        env.flush.fill_(1)                 # evict weights / state from L2 between timed steps (not timed)
        a.record(stream)
        _lib.check(lib.jlm_batch_run(batch))
        # the n-best fetch belongs to the step: it is where the near-tie guard re-scores / re-decodes in float64
        _lib.check(lib.jlm_batch_fetch(batch, C.byref(nb)))
        b.record(stream)
        _lib.check(lib.jlm_batch_get_info(batch, C.byref(info)))
        flagged += info.n_guard_flagged
        pairs += info.n_guard_pairs
        rerun += info.n_guard_rerun
    env.barrier()
    wall = time.perf_counter() - wall0
    ms_dev = env.reduce(sum(a.elapsed_time(b) for a, b in ev), 'MAX')
    total_chars = env.reduce(float(chars), 'SUM')
    value = total_chars * steps / (ms_dev * 1e-3)
    launches = int(info.kernel_launches) * steps
    rows_stepped = int(info.n_slots)
    guard = {'eps': float(info.guard_eps), 'flagged_fraction': flagged / float(steps * S),
             'flagged_sentences_per_step': flagged / float(steps), 'pairs_rescored_f64_per_step': pairs / float(steps),
             'sentences_redecoded_f64_per_step': rerun / float(steps), 'min_gap': float(info.guard_min_gap),
             'note': 'rank decisions resting on a score gap < eps are re-scored in float64 (jlm_pool); contradicted ones '
                     're-decode their sentence on the float64 back end; all inside the timed region'}

    # ---------------- roofline pass: same K steps with per-kernel CUDA events ----------------
    _lib.check(lib.jlm_batch_enable_timers(batch, 1))
    gate_ms = proj_ms = 0.0
    n_gate = n_proj = 0
    for _ in range(steps):
        env.flush.fill_(1)
        _lib.check(lib.jlm_batch_run(batch))
        _lib.check(lib.jlm_batch_fetch(batch, C.byref(nb)))
        _lib.check(lib.jlm_batch_get_info(batch, C.byref(info)))
        gate_ms += info.ms_gate_gemm
        proj_ms += info.ms_proj_gemm
        n_gate += info.n_gate_launches
        n_proj += info.n_proj_launches
    _lib.check(lib.jlm_batch_destroy(batch))

    # ---------------- the same K steps with the near-tie guard switched off (what the certification costs) ----------------
    if float(guard['eps']) > 0.0:
        _lib.check(lib.jlm_set_guard(hdl, 0.0))
        b3 = C.c_void_p()
        _lib.check(lib.jlm_batch_upload(hdl, C.byref(lb), BEAM, TOPN, MODE, args.backend, C.byref(b3)))
        _lib.check(lib.jlm_set_guard(hdl, -1.0))
        nb_off = (np.empty((S, TOPN)), np.empty(S, dtype=np.int32), np.empty((S, TOPN), dtype=np.int32),
                  np.zeros((S, TOPN, max_len), dtype=np.int32))
        nbo = _lib.NBest()
        nbo.top_n, nbo.max_len = TOPN, max_len
        nbo.scores, nbo.n_paths = _lib.ptr(nb_off[0], C.c_double), _lib.ptr(nb_off[1], C.c_int32)
        nbo.path_len, nbo.path_nodes = _lib.ptr(nb_off[2], C.c_int32), _lib.ptr(nb_off[3], C.c_int32)
        _lib.check(lib.jlm_batch_run(b3))
        _lib.check(lib.jlm_batch_fetch(b3, C.byref(nbo)))
        ev3 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        env.barrier()
        for a, b in ev3:
            env.flush.fill_(1)
            a.record(stream)
            _lib.check(lib.jlm_batch_run(b3))
            _lib.check(lib.jlm_batch_fetch(b3, C.byref(nbo)))
            b.record(stream)
        env.barrier()
        ms_off = env.reduce(sum(a.elapsed_time(b) for a, b in ev3), 'MAX')
        _lib.check(lib.jlm_batch_destroy(b3))
        guard['value_guard_off'] = total_chars * steps / (ms_off * 1e-3)
        guard['ms_per_step_guard_off'] = ms_off / steps
        guard['nbest_changed_by_guard'] = int(S - sum(
            int(n_paths[s] == nb_off[1][s] and np.array_equal(path_len[s], nb_off[2][s]) and
                np.array_equal(path_nodes[s], nb_off[3][s])) for s in range(S)))

    # ---------------- e2e arm: the public C-ABI calls with HOST buffers ----------------
    # kana text (UTF-32, host) -> jlm_lattice_build (host C++) -> plan, H2D, all frames, D2H -> n-best (host).
    # Nothing is resident on the device when the timed region starts except the model weights.
    lens = np.array([len(t) for t in sents], dtype=np.int64)
    tptr = np.zeros(len(sents) + 1, dtype=np.int64)
    np.cumsum(lens, out=tptr[1:])
    cps = np.frombuffer(''.join(sents).encode('utf-32-le'), dtype=np.uint32)
    tmax = int(lens.max()) + 1
    t_scores = np.empty((S, TOPN))
    t_npaths = np.empty(S, dtype=np.int32)
    t_len = np.empty((S, TOPN), dtype=np.int32)
    t_entry = np.zeros((S, TOPN, tmax), dtype=np.int32)
    t_start = np.zeros((S, TOPN, tmax), dtype=np.int32)
    tnb = _lib.TextNBest()
    tnb.top_n, tnb.max_len = TOPN, tmax
    tnb.scores, tnb.n_paths = _lib.ptr(t_scores, C.c_double), _lib.ptr(t_npaths, C.c_int32)
    tnb.path_len = _lib.ptr(t_len, C.c_int32)
    tnb.path_entry, tnb.path_start = _lib.ptr(t_entry, C.c_int32), _lib.ptr(t_start, C.c_int32)

    def e2e_submit():
        job = C.c_void_p()
        _lib.check(lib.jlm_decode_texts_submit(hdl, nlex.handle, len(sents), _lib.ptr(tptr, C.c_int64),
                                               _lib.ptr(cps, C.c_uint32), BEAM, TOPN, MODE, n_extra,
                                               _lib.ptr(extra, C.c_int32) if n_extra else None, args.backend,
                                               args.chunks, 0, C.byref(job)))
        return job

    def e2e_collect(job):
        _lib.check(lib.jlm_decode_texts_collect(job, C.byref(tnb), None))

    def e2e_run(n_steps, depth):
        """n_steps batches through submit/collect with at most `depth` in flight (depth 1 = the blocking
        jlm_decode_texts call); every batch is built from the host text again, copied H2D, decoded, copied
        D2H and unpacked inside the timed region."""
        env.barrier()
        t0 = time.perf_counter()
        inflight = []
        t_sub = t_col = 0.0
        for _ in range(n_steps):
            t1 = time.perf_counter()
            inflight.append(e2e_submit())
            t2 = time.perf_counter()
            if len(inflight) >= depth:
                e2e_collect(inflight.pop(0))
            t_sub += t2 - t1
            t_col += time.perf_counter() - t2
        while inflight:
            e2e_collect(inflight.pop(0))
        env.barrier()
        if os.environ.get('JLM_BENCH_E2E_DEBUG'):
            sys.stderr.write('[bench] e2e depth %d: %.2f ms per batch in submit, %.2f ms in collect (host thread)\n'
                             % (depth, 1e3 * t_sub / n_steps, 1e3 * t_col / n_steps))
        return env.reduce(time.perf_counter() - t0, 'MAX')

    for _ in range(2):
        e2e_collect(e2e_submit())
    # the e2e arm must reproduce the device-resident arm: same scores, same paths (as lexicon entries)
    assert np.array_equal(t_scores, scores) and np.array_equal(t_len, path_len), 'e2e arm and device-resident arm disagree'
    for s_i in (0, S // 2, S - 1):
        ids = path_nodes[s_i, 0, :path_len[s_i, 0]]
        assert np.array_equal(packed.node_entry[ids], t_entry[s_i, 0, :t_len[s_i, 0]]), 'e2e paths differ'
    e2e_run(args.e2e_depth + 2, args.e2e_depth)      # untimed warm-up: the second in-flight arena is allocated here
    e2e_s = e2e_run(steps, args.e2e_depth)
    assert np.array_equal(t_scores, scores) and np.array_equal(t_len, path_len), 'streamed e2e arm disagrees'
    e2e_blocking_s = e2e_run(steps, 1)
    e2e = total_chars * steps / e2e_s
    e2e_blocking = total_chars * steps / e2e_blocking_s
    # one more upload to read the byte counters of a single call
    b2 = C.c_void_p()
    _lib.check(lib.jlm_batch_upload(hdl, C.byref(lb), BEAM, TOPN, MODE, args.backend, C.byref(b2)))
    _lib.check(lib.jlm_batch_run(b2))
    _lib.check(lib.jlm_batch_fetch(b2, C.byref(nb)))
    i2 = _lib.BatchInfo()
    _lib.check(lib.jlm_batch_get_info(b2, C.byref(i2)))
    _lib.check(lib.jlm_batch_destroy(b2))

    # ---------------- strong scaling: ONE fixed set sharded over the ranks (shard.decode_sharded) ----------------
    strong = None
    if args.scaling == 'strong' or (not headline and name in ('cfg4', 'cfg5')) or (headline and world > 1):
        from jlm_b200 import synth
        fixed = synth.make_sentences(lexicon, args.sentences, min_len=MIN_LEN, seed=4242, vocab_size=wl['V'])
        kw = dict(topN=TOPN, beam_width=BEAM, backend=args.backend)
        if wl['dynamic']:
            kw.update(vocab_select=True, samples=wl['samples'], top_sampling=True)
        # the gathered n-best block stays in its array form (scores + lexicon-entry paths for all sentences on every
        # rank); building Python word lists from it is the caller's choice and is checked once, untimed
        full = shard.decode_sharded(dec, fixed, rank=rank, world_size=world, gather=True, **kw)      # warm-up, as words
        assert len(full) == len(fixed) and all(r is not None for r in full)
        k_strong = max(2, min(steps, 5))
        env.barrier()
        t0 = time.perf_counter()
        for _ in range(k_strong):
            arr = shard.decode_sharded(dec, fixed, rank=rank, world_size=world, gather=True, as_arrays=True, **kw)
        env.barrier()
        dt = env.reduce(time.perf_counter() - t0, 'MAX')
        assert arr['scores'].shape[0] == len(fixed)
        again = dec.words_from_arrays(fixed[:8], {k: v[:8] for k, v in arr.items()}, TOPN)
        assert [[w for _, w in r] for r in again] == [[w for _, w in r] for r in full[:8]]
        strong = {'value': sum(len(t) for t in fixed) * k_strong / dt, 'unit': 'chars/s', 'scaling': 'strong',
                  'sentences_total': len(fixed), 'sentences_per_gpu': int(np.ceil(len(fixed) / float(world))),
                  'steps': k_strong, 'ms_per_step': 1e3 * dt / k_strong,
                  'call': 'jlm_b200.shard.decode_sharded(as_arrays=True): length-balanced partition, host text -> n-best per '
                          'rank, packed all_gather_into_tensor (NCCL) of the n-best arrays, every rank ends with the whole '
                          'block in input order; wall clock, max over ranks'}
    t_load1 = time.perf_counter()
    env.windows.append((t_load0, t_load1))

    # ---------------- latency mode: one sentence per call (few rows in flight, exact float64 back end) ----------------
    lat_ms = lat_roof = None
    if rank == 0 and headline:
        one = lattice.NativeLattices(nlex, sents[:1], MODE, extra[:1] if extra is not None else None)
        v1 = one.c_struct()
        for _ in range(3):
            _lib.check(lib.jlm_decode_batch(hdl, C.byref(v1), BEAM, TOPN, MODE, _lib.BACKEND_AUTO, C.byref(nb)))
        t0 = time.perf_counter()
        for _ in range(20):
            _lib.check(lib.jlm_decode_batch(hdl, C.byref(v1), BEAM, TOPN, MODE, _lib.BACKEND_AUTO, C.byref(nb)))
        lat_ms = (time.perf_counter() - t0) / 20 * 1e3
        # roofline of the latency path's dominant kernel: the float64 weight stream over the output block(s)
        # (k_skinny_f64; HBM/L2-bound, algorithmic bytes = V * K * 4 per stepped frame), from the CUDA-event
        # "softmax" bucket of one timed call (that bucket also holds the LSE merge and the needed-word dots)
        b1 = C.c_void_p()
        _lib.check(lib.jlm_batch_upload(hdl, C.byref(v1), BEAM, TOPN, MODE, _lib.BACKEND_AUTO, C.byref(b1)))
        _lib.check(lib.jlm_batch_enable_timers(b1, 1))
        _lib.check(lib.jlm_batch_run(b1))
        nb1 = _lib.NBest()
        s1 = (np.empty((1, TOPN)), np.empty(1, dtype=np.int32), np.empty((1, TOPN), dtype=np.int32),
              np.zeros((1, TOPN, max_len), dtype=np.int32))
        nb1.top_n, nb1.max_len = TOPN, max_len
        nb1.scores, nb1.n_paths = _lib.ptr(s1[0], C.c_double), _lib.ptr(s1[1], C.c_int32)
        nb1.path_len, nb1.path_nodes = _lib.ptr(s1[2], C.c_int32), _lib.ptr(s1[3], C.c_int32)
        _lib.check(lib.jlm_batch_fetch(b1, C.byref(nb1)))
        i1 = _lib.BatchInfo()
        _lib.check(lib.jlm_batch_get_info(b1, C.byref(i1)))
        _lib.check(lib.jlm_batch_destroy(b1))
        if wl['mode'] == 'dsoftmax_star':
            out_bytes = sum(4.0 * sz * ((wl['V'] if e is None else e) - st) for sz, st, e in wl['segments'])
        else:
            out_bytes = 4.0 * wl['E'] * wl['V']
        n_frames = len(sents[0]) + (0 if wl['dynamic'] else 1)
        lat_roof = None
        if not wl['dynamic'] and i1.ms_softmax > 0:
            # One sentence alone runs in ONE cooperative kernel (k_single_f64) unless per-bucket timers are on, so the
            # roofline is taken over the whole call: every frame streams the output block(s), the gate weights and the
            # stage-1 matrix once (all L2-resident between frames).  The per-bucket CUDA-event times of the per-frame
            # launch path (169 launches, timers on) are kept beside it.
            gate_bytes = 4.0 * 4 * wl['H'] * (wl['H'] + wl['E'])
            stage1_bytes = 8.0 * wl['H'] * (sum(sz for sz, _, _ in wl['segments']) if wl['mode'] == 'dsoftmax_star' else wl['E'])
            frame_bytes = out_bytes + gate_bytes + stage1_bytes
            gbs = frame_bytes * n_frames / (lat_ms * 1e-3) / 1e9
            lat_roof = {'bound': 'hbm', 'kernel': 'k_single_f64 (cooperative kernel, whole frame loop of one sentence: float64 '
                                                   'weight streams of the gate, stage-1 and output layers, <= 16 rows)',
                        'achieved': gbs, 'peak': peaks()[2], 'unit': 'GB/s', 'frac': gbs / peaks()[2],
                        'bytes_per_frame': frame_bytes, 'frames': n_frames, 'ms_call': lat_ms,
                        'launch_path_ms_softmax_bucket': float(i1.ms_softmax),
                        'launch_path_ms_lstm_bucket': float(i1.ms_lstm), 'launch_path_ms_beam_bucket': float(i1.ms_beam),
                        'note': 'weights are L2-resident between frames (58 MB < 126 MB): an L2 stream; the kernel is bound by '
                                'float64 FMA issue and shared-memory broadcasts, not by bandwidth (profiles/r02/ncu_single.csv); '
                                'ms_call = upload + kernel + n-best fetch of one jlm_decode_batch call'}
        _lib.check(lib.jlm_decode_batch(hdl, C.byref(lb), BEAM, TOPN, MODE, args.backend, C.byref(nb)))   # restore nb

    if rank != 0:
        del dec
        return None

    sus, burst, hbm, src = peaks()
    f_gate_row, f_proj_row = flops_per_row(wl)
    rows_total = rows_stepped * steps
    traffic = None
    tp = os.path.join(REPO, 'profiles', 'traffic.json')
    if os.path.exists(tp) and name == 'cfg2':       # dram bytes per launch from the committed ncu capture
        traffic = json.load(open(tp)).get('k_tc_gemm<256,EPI_LSE>', {}).get('dram_bytes_per_launch')
    roof = {'bound': 'tensor', 'unit': 'TFLOP/s', 'peak': sus, 'peak_source': src + ' bf16 sustained', 'traffic': traffic}
    if proj_ms > 0 and args.backend == 2:
        ach = f_proj_row * rows_total / (proj_ms * 1e-3) / 1e12
        if wl['mode'] == 'dsoftmax_star':
            # one launch per segment and frame: K <= 128 -> row-stationary k_tc_lse_rs<1|2>, K = 256 -> k_tc_gemm<256,EPI_LSE>
            kname = ('k_tc_lse_rs<K/64> (row-stationary, K <= 128) + k_tc_gemm<256,EPI_LSE,cta_group::2> (K = 256): output '
                     'projection segments + online LSE, all launches of a frame')
        else:
            kname = 'k_tc_gemm<256,EPI_LSE,cta_group::2> (output projection + online LSE)'
        roof.update({'kernel': kname, 'achieved': ach,
                     'frac': ach / sus, 'issued_frac': 3 * ach / sus,
                     'flops_per_launch': f_proj_row * rows_total / max(n_proj, 1),
                     'avg_launch_ms': proj_ms / max(n_proj, 1), 'launches': n_proj,
                     'note': 'useful flops; each product is 3 fp16 MMAs (2-term split), issued_frac counts those'})
        g_ach = f_gate_row * rows_total / (gate_ms * 1e-3) / 1e12 if gate_ms > 0 else None
        roof['gate'] = {'kernel': 'k_tc_gemm<256,EPI_LSTM,cta_group::2> (gate GEMM + LSTM epilogue)', 'achieved': g_ach,
                        'frac': g_ach / sus if g_ach else None, 'issued_frac': 3 * g_ach / sus if g_ach else None,
                        'avg_launch_ms': gate_ms / max(n_gate, 1), 'launches': n_gate}
    elif gate_ms > 0 and args.backend == 2:
        # vocabulary-selection workloads have no full-vocabulary GEMM: the gate GEMM is the tensor kernel
        g_ach = f_gate_row * rows_total / (gate_ms * 1e-3) / 1e12
        roof.update({'kernel': 'k_tc_gemm<256,EPI_LSTM,cta_group::2> (gate GEMM + LSTM epilogue)', 'achieved': g_ach,
                     'frac': g_ach / sus, 'issued_frac': 3 * g_ach / sus, 'traffic': None,
                     'flops_per_launch': f_gate_row * rows_total / max(n_gate, 1),
                     'avg_launch_ms': gate_ms / max(n_gate, 1), 'launches': n_gate})
    else:
        roof.update({'achieved': None, 'frac': None})

    # ---------------- CPU baseline: oracle port on a bounded sample of the same workload ----------------
    nb_cpu = max(1, min(cpu_sentences if world == 1 else min(cpu_sentences, 2), len(sents)))
    try:      # all host cores for the CPU leg, whatever thread limit the launcher exported
        import threadpoolctl
        threadpoolctl.threadpool_limits(limits=os.cpu_count())
    except Exception:
        pass
    prepare, run_cpu = oracle_decoder(wl, cfg, weights, lexicon, reading_dict)
    sub = sents[:nb_cpu]
    states = [prepare(s) for s in sub]
    t0 = time.perf_counter()
    cpu_out = [run_cpu(st) for st in states]
    cpu_s = time.perf_counter() - t0
    cpu_chars = sum(len(s) for s in sub)
    # parity spot-check of the timed GPU output against the oracle on the same sentences
    top1_same = nbest_same = 0
    for s in range(nb_cpu):
        got = [[w for w in packed.path_words(s, path_nodes[s, k, :path_len[s, k]].tolist()) if w != '<eos>']
               for k in range(int(n_paths[s]))]
        top1_same += int(got[0] == cpu_out[s][0][1])
        nbest_same += int(got == [ws for _, ws in cpu_out[s]])
    del dec

    line = {
        'metric': METRIC, 'value': value, 'unit': 'chars/s', 'n_gpus': world, 'steps': steps,
        'warmup': max(warmup, 3), 'ms_per_step': ms_dev / steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'fp16x2-split->f32' if args.backend == 2 else 'f64',
        'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'sentences_per_gpu_per_step': S, 'chars_per_gpu_per_step': chars,
                   'lm_rows_per_step': rows_stepped, 'backend': 'tcgen05' if args.backend == 2 else 'exact-f64',
                   'l2': 'explicit 256 MiB flush between timed steps', 'host_lattice_build_s': t_lat,
                   'single_sentence_latency_ms': lat_ms, 'single_sentence_chars': len(sents[0]),
                   'single_sentence_roofline': lat_roof,
                   'wall_s_timed_region': wall,
                   'timed_region': 'jlm_batch_run + jlm_batch_fetch per step (frames, n-best D2H, near-tie guard)'},
        'e2e': {'value': e2e, 'unit': 'chars/s', 'h2d_bytes_per_step': int(i2.h2d_bytes),
                'd2h_bytes_per_step': int(i2.d2h_bytes),
                'call': 'jlm_decode_texts_submit + jlm_decode_texts_collect, %d batches in flight (host UTF-32 kana -> '
                        'host n-best paths; per batch: lattice build, plan, H2D, frames, D2H, guard, unpack - all inside the '
                        'timed region; the host work of batch k+1 overlaps the device work of batch k)' % args.e2e_depth,
                'blocking_value': e2e_blocking,
                'blocking_call': 'jlm_decode_texts (one blocking call per batch, nothing overlapped)'},
        'gpu_launches': launches,
        'roofline': roof,
        'guard': guard,
        'cpu_baseline': {'value': cpu_chars / cpu_s if world == 1 else None, 'unit': 'chars/s', 'cores': os.cpu_count(),
                         'kind': 'port', 'note': CPU_NOTE if world == 1 else 'timed at N=1 only; parity spot-check kept',
                         'sample': 'first %d sentences (%d chars), lattice->n-best, numpy oracle' % (nb_cpu, cpu_chars),
                         'top1_identical_to_gpu': '%d/%d' % (top1_same, nb_cpu),
                         'nbest_identical_to_gpu': '%d/%d' % (nbest_same, nb_cpu)},
    }
    if strong is not None:
        line['strong'] = strong
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--sentences', type=int, default=1024, help='sentences per GPU per step (lock-step batch)')
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--ref-sentences', type=int, default=8, help='reference arm: sentences decoded per step')
    ap.add_argument('--cpu-baseline-sentences', type=int, default=32)
    ap.add_argument('--backend', type=int, default=2, help='1 exact (float64 CUDA cores), 2 tensor cores')
    ap.add_argument('--workload', default='cfg2', choices=sorted(WORKLOADS))
    ap.add_argument('--extra', default='cfg3,cfg4,cfg5',
                    help='other BASELINE configs run for --extra-steps each and reported under "workloads" (none = skip)')
    ap.add_argument('--extra-steps', type=int, default=3)
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                    help='strong: also time the headline workload as ONE fixed set sharded over the ranks')
    ap.add_argument('--chunks', type=int, default=0, help='e2e arm: pipeline chunks of jlm_decode_texts (0 = automatic)')
    ap.add_argument('--e2e-depth', type=int, default=3, help='e2e arm: batches in flight through submit/collect')
    ap.add_argument('--profile', action='store_true', help='1 warm-up + K plain steps only (for ncu); prints no JSON')
    args = ap.parse_args()

    if args.impl == 'reference':
        run_reference(args, int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1')))
        return

    env = Env(args)
    line = bench_workload(env, args.workload, args.steps, args.warmup, args.cpu_baseline_sentences, headline=True)
    extras = []
    if not args.profile and args.extra and args.extra != 'none':
        for name in [w for w in args.extra.split(',') if w in WORKLOADS and w != args.workload]:
            sub = bench_workload(env, name, args.extra_steps, 3, 2 if name != 'cfg5' else 1, headline=False)
            if sub is not None:
                keep = ('value', 'unit', 'steps', 'ms_per_step', 'dtype', 'e2e', 'gpu_launches', 'roofline', 'guard',
                        'cpu_baseline', 'strong')
                extras.append(dict({'workload': name, 'config': sub['config']}, **{k: sub[k] for k in keep if k in sub}))
    if env.rank == 0 and line is not None:
        line['workloads'] = extras
        line['clocks'] = env.sampler.stop(env.windows)
        print(json.dumps(line), flush=True)
    if env.world > 1:
        env.dist.destroy_process_group()


if __name__ == '__main__':
    main()
