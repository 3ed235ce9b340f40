"""Host-side batch planning through the C ABI (``oe_plan_speeds`` / ``oe_plan_augment``): the reference's
random decisions for a whole batch in two native calls instead of ~10 Python ``random`` calls per
utterance.  The native generator continues Python's global ``random`` state (Mersenne Twister, CPython's
``random()`` / ``_randbelow`` semantics) and the advanced state is written back, so results -- and any
later ``random`` use by the caller -- are identical to running the reference's Python loops
(tests/test_host_logic.py::test_native_planner_continues_python_random)."""
import ctypes
import random

import numpy as np

from . import _lib
from ._lib import c_f64p, c_i32p, c_u32p, c_u8p, check


def _get_state():
    return np.array(random.getstate()[1], dtype=np.uint32)


def _set_state(st):
    random.setstate((3, tuple(st.tolist()), None))


def plan_speeds(perturb_rate, speeds_cfg, item_speeds, active=None):
    """dataset.py:87-89 + audio_processor.py:5-18 for every utterance, in input order -> float64 [n]."""
    lib = _lib.load()
    item = np.ascontiguousarray(item_speeds, dtype=np.float64)
    n = item.shape[0]
    out = np.empty(n, dtype=np.float64)
    cfg = None if speeds_cfg is None else np.ascontiguousarray([float(s) for s in speeds_cfg], dtype=np.float64)
    act = None if active is None else np.ascontiguousarray(active, dtype=np.uint8)
    st = _get_state()
    check(lib.oe_plan_speeds(st.ctypes.data_as(c_u32p), n, float(perturb_rate),
                             cfg.ctypes.data_as(c_f64p) if cfg is not None else None,
                             0 if cfg is None else cfg.shape[0], item.ctypes.data_as(c_f64p),
                             act.ctypes.data_as(c_u8p) if act is not None else None, out.ctypes.data_as(c_f64p)))
    _set_state(st)
    return out


def plan_augment(frames, num_freq, spec_sub_conf=None, spec_aug_conf=None):
    """dataset.py:204-209 for a length-sorted batch: all _spec_substitute draws, then all
    _spec_augmentation draws.  Returns (frame_map or None [sum frames], tmask or None [n, k, 2],
    fmask or None [n, k, 2])."""
    lib = _lib.load()
    frames = np.ascontiguousarray(frames, dtype=np.int32)
    n = frames.shape[0]
    do_sub = spec_sub_conf is not None
    do_aug = spec_aug_conf is not None
    sub = dict(max_t=20, num_t_sub=3)                      # feature_processor.py:44 defaults
    sub.update(spec_sub_conf or {})
    aug = dict(num_t_mask=2, num_f_mask=2, max_t=50, max_f=10)   # feature_processor.py:10-14 defaults
    aug.update(spec_aug_conf or {})
    fmap = np.empty(int(frames.sum()) if do_sub else 0, dtype=np.int32)
    tm = np.empty((n, aug['num_t_mask'], 2) if do_aug else (0, 0, 2), dtype=np.int32)
    fm = np.empty((n, aug['num_f_mask'], 2) if do_aug else (0, 0, 2), dtype=np.int32)
    if n == 0 or not (do_sub or do_aug):
        return (fmap if do_sub else None), None, None
    st = _get_state()
    check(lib.oe_plan_augment(st.ctypes.data_as(c_u32p), n, frames.ctypes.data_as(c_i32p), int(num_freq),
                              int(do_sub), int(sub['max_t']), int(sub['num_t_sub']), int(do_aug),
                              int(aug['num_t_mask']), int(aug['num_f_mask']), int(aug['max_t']), int(aug['max_f']),
                              fmap.ctypes.data_as(c_i32p), tm.ctypes.data_as(c_i32p), fm.ctypes.data_as(c_i32p)))
    _set_state(st)
    return (fmap if do_sub else None), (tm if do_aug and tm.shape[1] else None), (fm if do_aug and fm.shape[1] else None)
