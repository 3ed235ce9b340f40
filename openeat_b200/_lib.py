"""ctypes binding of ``include/openeat_frontend.h`` (the C ABI of the CUDA library).

The product path has NO CPU fallback: if ``libopeneat_frontend.so`` is missing or a
call fails, an exception is raised.  ``build()`` compiles the library in-tree with
nvcc for sm_100a (it cross-compiles without a GPU).
"""
import ctypes
import os
import subprocess

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'csrc')
LIB_PATH = os.environ.get('OE_LIB_PATH') or os.path.join(CSRC, 'libopeneat_frontend.so')   # OE_LIB_PATH: developer A/B builds
EMUL_PATH = os.path.join(CSRC, 'liboe_emul.so')
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '--shared', '-Xcompiler', '-fPIC', '-diag-suppress', '177,550']

OE_OK = 0
OE_ERR_INVALID, OE_ERR_UNSUPPORTED, OE_ERR_CUDA, OE_ERR_WORKSPACE = 1, 2, 3, 4
OE_WAV_I16, OE_WAV_F32, OE_FEATS_F32 = 0, 1, 2
OE_NORM_NONE, OE_NORM_PER_UTT = 0, 1

c_i32p = ctypes.POINTER(ctypes.c_int32)
c_i64p = ctypes.POINTER(ctypes.c_int64)
c_f32p = ctypes.POINTER(ctypes.c_float)
c_f64p = ctypes.POINTER(ctypes.c_double)
c_u32p = ctypes.POINTER(ctypes.c_uint32)
c_u8p = ctypes.POINTER(ctypes.c_uint8)


class OeConfig(ctypes.Structure):
    _fields_ = [('sample_rate', ctypes.c_int32), ('frame_length', ctypes.c_int32),
                ('frame_shift', ctypes.c_int32), ('fft_size', ctypes.c_int32),
                ('num_mel_bins', ctypes.c_int32), ('preemph', ctypes.c_float),
                ('low_freq', ctypes.c_float), ('high_freq', ctypes.c_float),
                ('log_floor', ctypes.c_float)]


class OeBatch(ctypes.Structure):
    _fields_ = [('batch', ctypes.c_int32), ('wav_dtype', ctypes.c_int32),
                ('wav_offsets', c_i64p), ('wav_lens', c_i32p),
                ('out_rows', c_i64p), ('out_nrows', c_i32p),
                ('out_pitch', ctypes.c_int64), ('norm_mode', ctypes.c_int32),
                ('n_tmask', ctypes.c_int32), ('n_fmask', ctypes.c_int32),
                ('tmask', c_i32p), ('fmask', c_i32p),
                ('frame_map', c_i32p), ('frame_map_offsets', c_i64p),
                ('d_cmvn_mean', ctypes.c_void_p), ('d_cmvn_istd', ctypes.c_void_p),
                ('cmvn_on_padding', ctypes.c_int32), ('d_stats', ctypes.c_void_p),
                ('out_frames', c_i32p), ('resample_ids', c_i32p),
                ('feature_dither', ctypes.c_float), ('dither_seed', ctypes.c_uint64),
                ('wav_dither', ctypes.c_float)]


class OeResampleBatch(ctypes.Structure):
    _fields_ = [('batch', ctypes.c_int32), ('wav_dtype', ctypes.c_int32),
                ('in_offsets', c_i64p), ('in_lens', c_i32p), ('table_ids', c_i32p),
                ('out_offsets', c_i64p), ('out_lens', c_i32p), ('orig_rates', c_i32p), ('new_rates', c_i32p)]


OE_RS_DIRECT = -2


# every symbol include/openeat_frontend.h declares: (restype, argtypes)
SYMBOLS = {
    'oe_last_error': (ctypes.c_char_p, []),
    'oe_abi_version': (ctypes.c_int, []),
    'oe_config_default': (ctypes.c_int, [ctypes.POINTER(OeConfig)]),
    'oe_frontend_create': (ctypes.c_int, [ctypes.POINTER(OeConfig), c_f32p, c_f32p, ctypes.c_int,
                                          ctypes.POINTER(ctypes.c_void_p)]),
    'oe_frontend_destroy': (ctypes.c_int, [ctypes.c_void_p]),
    'oe_frontend_launch_count': (ctypes.c_int64, [ctypes.c_void_p]),
    'oe_frontend_set_kernel_timing': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32]),
    'oe_frontend_fbank_kernel_ms': (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_float)]),
    'oe_frontend_step_ms': (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_float)]),
    'oe_frontend_get_tables': (ctypes.c_int, [ctypes.c_void_p, c_f32p, c_f32p]),
    'oe_num_frames': (ctypes.c_int32, [ctypes.c_void_p, ctypes.c_int64]),
    'oe_fbank_workspace_bytes': (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(OeBatch),
                                                ctypes.POINTER(ctypes.c_size_t)]),
    'oe_fbank_batch': (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(OeBatch), ctypes.c_void_p,
                                      ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    'oe_batch_prepare': (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(OeBatch), ctypes.POINTER(ctypes.c_void_p)]),
    'oe_prepared_destroy': (ctypes.c_int, [ctypes.c_void_p]),
    'oe_prepared_workspace_bytes': (ctypes.c_size_t, [ctypes.c_void_p]),
    'oe_prepared_frames': (c_i32p, [ctypes.c_void_p]),
    'oe_fbank_run': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.c_size_t, ctypes.c_void_p]),
    'oe_upload_small': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]),
    'oe_cmvn_apply': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32,
                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    'oe_cmvn_conv_subsample': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                              ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32,
                                              ctypes.c_void_p, ctypes.c_void_p]),
    'oe_add_resampler': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, c_f32p,
                                        ctypes.c_int32, c_i32p]),
    'oe_resampler_fusable': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32]),
    'oe_mel_is_baked': (ctypes.c_int, [ctypes.c_void_p]),
    'oe_resample_out_len': (ctypes.c_int64, [ctypes.c_int64, ctypes.c_int32, ctypes.c_int32]),
    'oe_resample_workspace_bytes': (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(OeResampleBatch),
                                                   ctypes.POINTER(ctypes.c_size_t)]),
    'oe_resample': (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(OeResampleBatch), ctypes.c_void_p,
                                   ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    'oe_ingest_create': (ctypes.c_int, [ctypes.c_int32, ctypes.POINTER(ctypes.c_void_p)]),
    'oe_ingest_destroy': (ctypes.c_int, [ctypes.c_void_p]),
    'oe_ingest_probe': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.POINTER(ctypes.c_char_p), c_f64p, c_f64p,
                                       c_i32p, c_i32p, c_i32p]),
    'oe_ingest_read': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.POINTER(ctypes.c_char_p), c_f64p, c_f64p,
                                      ctypes.c_void_p, c_i64p, c_i32p, c_i32p]),
    'oe_ingest_error': (ctypes.c_char_p, [ctypes.c_void_p, ctypes.c_int32]),
    'oe_ingest_submit': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.POINTER(ctypes.c_char_p), c_f64p, c_f64p,
                                        ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_void_p)]),
    'oe_ingest_wait': (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(c_i64p), ctypes.POINTER(c_i32p), ctypes.POINTER(c_i32p),
                                      ctypes.POINTER(c_i32p), ctypes.POINTER(ctypes.c_int64)]),
    'oe_ingest_job_error': (ctypes.c_char_p, [ctypes.c_void_p, ctypes.c_int32]),
    'oe_ingest_job_release': (ctypes.c_int, [ctypes.c_void_p]),
    'oe_host_pad_rows': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, c_i32p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                        ctypes.c_void_p, ctypes.c_void_p]),
    'oe_flac_info': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, c_i32p, c_i32p, c_i32p, c_i64p]),
    'oe_flac_decode': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int64, ctypes.c_int64,
                                      ctypes.c_void_p, ctypes.c_int32, c_i64p]),
    'oe_flac_pack': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.POINTER(ctypes.c_char_p), c_f64p, c_f64p,
                                    ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, c_i64p, c_i64p, c_i32p, c_i32p,
                                    c_i32p, c_i64p, c_i64p, c_i64p]),
    'oe_flac_submit': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.POINTER(ctypes.c_char_p), c_f64p, c_f64p,
                                      ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_void_p)]),
    'oe_flac_wait': (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(c_i64p), ctypes.POINTER(c_i32p), ctypes.POINTER(c_i32p),
                                    ctypes.POINTER(c_i32p), c_i64p, c_i64p, c_i64p]),
    'oe_flac_decode_batch': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                            ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p]),
    'oe_flac_encode': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                      ctypes.c_void_p, ctypes.c_int64, c_i64p]),
    'oe_plan_speeds': (ctypes.c_int, [c_u32p, ctypes.c_int32, ctypes.c_double, c_f64p, ctypes.c_int32, c_f64p, c_u8p,
                                      c_f64p]),
    'oe_plan_augment': (ctypes.c_int, [c_u32p, ctypes.c_int32, c_i32p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                       ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                       ctypes.c_int32, c_i32p, c_i32p, c_i32p]),
}

_lib = None


class OeFlacFrame(ctypes.Structure):
    """``oe_flac_frame`` (include/openeat_frontend.h): one audio frame of a packed FLAC batch."""
    _fields_ = [('comp_off', ctypes.c_int64), ('out_off', ctypes.c_int64), ('frame_bytes', ctypes.c_int32),
                ('hdr_bytes', ctypes.c_int32), ('block', ctypes.c_int32), ('bps', ctypes.c_int32), ('skip', ctypes.c_int32),
                ('take', ctypes.c_int32), ('utt', ctypes.c_int32), ('reserved', ctypes.c_int32)]


class FrontendError(RuntimeError):
    pass


def build(force=False, verbose=False):
    """nvcc -> csrc/libopeneat_frontend.so (sm_100a), g++ -> csrc/liboe_emul.so (test tooling)."""
    src = os.path.join(CSRC, 'oe_frontend.cu')
    deps = [src, os.path.join(CSRC, 'oe_fft.h'), os.path.join(CSRC, 'oe_ingest.h'), os.path.join(CSRC, 'oe_flac.h'), os.path.join(CSRC, 'oe_flac_gpu.cuh'), os.path.join(CSRC, 'oe_fbank_kernel.cuh'),
            os.path.join(CSRC, 'oe_fbank2_kernel.cuh'), os.path.join(CSRC, 'oe_mel80.h'),
            os.path.join(CSRC, 'oe_rs_coefs.h'),
            os.path.join(os.path.dirname(CSRC), '..', 'include', 'openeat_frontend.h')]
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(d) > os.path.getmtime(LIB_PATH) for d in deps):
        cmd = ['nvcc'] + NVCC_FLAGS + ['-o', LIB_PATH, src]
        if verbose:
            cmd.insert(1, '-Xptxas=-v')
        subprocess.run(cmd, check=True, cwd=CSRC)
    emul = os.path.join(CSRC, 'oe_emul.cpp')
    if force or not os.path.exists(EMUL_PATH) or os.path.getmtime(emul) > os.path.getmtime(EMUL_PATH) \
            or any(os.path.getmtime(os.path.join(CSRC, d)) > os.path.getmtime(EMUL_PATH) for d in ('oe_fft.h', 'oe_flac_gpu.cuh')):
        subprocess.run(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-o', EMUL_PATH, emul],
                       check=True, cwd=CSRC)
    return LIB_PATH


def load():
    """Loads the CUDA library; raises FrontendError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FrontendError('%s is missing: run `python -c "import __graft_entry__ as g; g.build()"` '
                            '(there is no CPU fallback for the front-end)' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != OE_OK:
        msg = load().oe_last_error()
        raise FrontendError('openeat_frontend error %d: %s' % (rc, msg.decode() if msg else '?'))
