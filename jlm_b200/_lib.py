"""ctypes binding of libjlm_b200.so (include/jlm_b200.h).  No CPU fallback: if the shared library
has not been built, or no CUDA device is usable, callers get a JlmError."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'csrc', 'libjlm_b200.so')

MAX_SEGMENTS = 8
MAX_BEAM = 128
BEAM_UNLIMITED = 0       # beam_width=None (decoder.py:227-229 skipped)
PROJ_UNTIED, PROJ_TIED, PROJ_DSOFTMAX, PROJ_DSOFTMAX_STAR = 0, 1, 2, 3
BACKEND_AUTO, BACKEND_EXACT, BACKEND_TC = 0, 1, 2
DECODE_FULL, DECODE_STATIC_VOCAB, DECODE_DYNAMIC = 0, 1, 2

_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_u32p = C.POINTER(C.c_uint32)


class Config(C.Structure):
    _fields_ = [('vocab_size', C.c_int32), ('hidden_size', C.c_int32), ('input_embed', C.c_int32),
                ('proj_mode', C.c_int32), ('self_norm', C.c_int32), ('n_seg', C.c_int32),
                ('seg_width', C.c_int32 * MAX_SEGMENTS), ('seg_start', C.c_int32 * MAX_SEGMENTS),
                ('seg_end', C.c_int32 * MAX_SEGMENTS)]


class Weights(C.Structure):
    _fields_ = [('HM', _f32p * 4), ('IM', _f32p * 4), ('b', _f32p * 4), ('b2', _f32p), ('LM_in', _f32p),
                ('PM', _f32p), ('UM', _f32p), ('seg_LM', _f32p * MAX_SEGMENTS), ('seg_VT', _f32p * MAX_SEGMENTS)]


class LatticeBatch(C.Structure):
    _fields_ = [('n_sent', C.c_int32), ('sent_len', _i32p), ('frame_ptr_off', _i64p), ('frame_ptr', _i64p),
                ('node_start', _i32p), ('node_word', _i32p), ('vocab_ptr', _i64p), ('vocab_ids', _i32p),
                ('vocab_frame_ptr', _i32p), ('dup_ptr', _i64p), ('dup_ids', _i32p)]


class NBest(C.Structure):
    _fields_ = [('top_n', C.c_int32), ('max_len', C.c_int32), ('scores', _f64p), ('n_paths', _i32p),
                ('path_len', _i32p), ('path_nodes', _i32p)]


class TextNBest(C.Structure):
    _fields_ = [('top_n', C.c_int32), ('max_len', C.c_int32), ('scores', _f64p), ('n_paths', _i32p),
                ('path_len', _i32p), ('path_entry', _i32p), ('path_start', _i32p)]


class BatchInfo(C.Structure):
    _fields_ = [('n_slots', C.c_int64), ('n_candidates', C.c_int64), ('n_nodes', C.c_int64),
                ('n_steps', C.c_int32), ('backend', C.c_int32), ('kernel_launches', C.c_int64),
                ('h2d_bytes', C.c_int64), ('d2h_bytes', C.c_int64), ('ms_lstm', C.c_float),
                ('ms_softmax', C.c_float), ('ms_beam', C.c_float), ('ms_gate_gemm', C.c_float),
                ('ms_proj_gemm', C.c_float), ('n_gate_launches', C.c_int32), ('n_proj_launches', C.c_int32),
                ('beam_width', C.c_int32), ('n_guard_flagged', C.c_int32), ('guard_min_gap', C.c_double),
                ('guard_eps', C.c_double), ('n_guard_pairs', C.c_int32), ('n_guard_rerun', C.c_int32)]


# every symbol include/jlm_b200.h declares: name -> (restype, argtypes)
_VP = C.c_void_p
SYMBOLS = {
    'jlm_last_error': (C.c_char_p, []),
    'jlm_abi_version': (C.c_int32, []),
    'jlm_create': (C.c_int32, [C.POINTER(Config), C.POINTER(Weights), C.c_int32, C.POINTER(_VP)]),
    'jlm_destroy': (C.c_int32, [_VP]),
    'jlm_set_quantized_block': (C.c_int32, [_VP, C.c_int32, C.POINTER(C.c_uint8), _f32p, C.c_int32]),
    'jlm_set_stream': (C.c_int32, [_VP, _VP]),
    'jlm_set_guard': (C.c_int32, [_VP, C.c_double]),
    'jlm_set_guard_verify': (C.c_int32, [_VP, C.c_int32]),
    'jlm_set_guard_scope': (C.c_int32, [_VP, C.c_int32]),
    'jlm_synchronize': (C.c_int32, [_VP]),
    'jlm_lstm_step': (C.c_int32, [_VP, _i32p, _f64p, _f64p, C.c_int32, _f64p, _f64p]),
    'jlm_project': (C.c_int32, [_VP, _f64p, C.c_int32, _i32p, _i32p, C.c_int32, _f64p]),
    'jlm_predict': (C.c_int32, [_VP, _i32p, _f64p, _f64p, C.c_int32, _i32p, _i32p, C.c_int32, _f64p, _f64p,
                                _f64p, _f64p, _f32p, _f32p]),
    'jlm_decode_batch': (C.c_int32, [_VP, C.POINTER(LatticeBatch), C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                     C.POINTER(NBest)]),
    'jlm_decode_texts': (C.c_int32, [_VP, _VP, C.c_int32, _i64p, _u32p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _i32p,
                                     C.c_int32, C.c_int32, C.POINTER(TextNBest), C.POINTER(BatchInfo)]),
    'jlm_decode_texts_submit': (C.c_int32, [_VP, _VP, C.c_int32, _i64p, _u32p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                            _i32p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(_VP)]),
    'jlm_decode_texts_collect': (C.c_int32, [_VP, C.POINTER(TextNBest), C.POINTER(BatchInfo)]),
    'jlm_decode_texts_cancel': (C.c_int32, [_VP]),
    'jlm_batch_fetch_async': (C.c_int32, [_VP]),
    'jlm_pool_create': (C.c_int32, [_VP, C.c_int64, C.POINTER(_VP)]),
    'jlm_pool_destroy': (C.c_int32, [_VP]),
    'jlm_pool_reset': (C.c_int32, [_VP]),
    'jlm_pool_step': (C.c_int32, [_VP, C.c_int32, _i32p, _i32p, C.POINTER(C.c_int64)]),
    'jlm_pool_nll': (C.c_int32, [_VP, C.c_int32, _i32p, _i32p, _f64p]),
    'jlm_pool_get_state': (C.c_int32, [_VP, C.c_int64, C.c_int32, _f64p, _f64p]),
    'jlm_batch_upload': (C.c_int32, [_VP, C.POINTER(LatticeBatch), C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                     C.POINTER(_VP)]),
    'jlm_batch_run': (C.c_int32, [_VP]),
    'jlm_batch_fetch': (C.c_int32, [_VP, C.POINTER(NBest)]),
    'jlm_batch_destroy': (C.c_int32, [_VP]),
    'jlm_batch_get_info': (C.c_int32, [_VP, C.POINTER(BatchInfo)]),
    'jlm_batch_enable_timers': (C.c_int32, [_VP, C.c_int32]),
    'jlm_batch_get_beams': (C.c_int32, [_VP, C.c_int32, _i32p, _f64p, _i32p, _i32p, _i32p, _f64p, _f64p, _f64p]),
    'jlm_lexicon_create': (C.c_int32, [C.c_int32, _i64p, _u32p, _i64p, _i32p, C.c_int32, C.c_int32, C.POINTER(_VP)]),
    'jlm_lexicon_destroy': (C.c_int32, [_VP]),
    'jlm_lattice_build': (C.c_int32, [_VP, C.c_int32, _i64p, _u32p, C.c_int32, C.c_int32, _i32p, C.POINTER(_VP)]),
    'jlm_lattice_view': (C.c_int32, [_VP, C.POINTER(LatticeBatch), C.POINTER(_i32p), C.POINTER(C.c_int64)]),
    'jlm_lattice_destroy': (C.c_int32, [_VP]),
    'jlm_tc_gemm_selftest': (C.c_int32, [_VP, _f32p, _f32p, C.c_int32, C.c_int32, C.c_int32, _f32p, _f32p]),
}

_lib = None


class JlmError(RuntimeError):
    pass


def load():
    """Loads the CUDA library (once).  Raises if it has not been built - there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise JlmError('libjlm_b200.so is not built (%s missing); run `python -c "import __graft_entry__ as g; '
                       'g.build()"` or `make -C jlm_b200/csrc`. There is no CPU fallback.' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.jlm_abi_version() != 1:
        raise JlmError('libjlm_b200.so ABI version mismatch')
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise JlmError(load().jlm_last_error().decode('utf-8', 'replace'))


def ptr(arr, ctype):
    """ctypes pointer to a contiguous numpy array (None -> NULL)."""
    if arr is None:
        return None
    return arr.ctypes.data_as(C.POINTER(ctype))
