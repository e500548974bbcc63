"""ctypes mirror of include/lgarb_group.h (liblgarb_group.so): several GPUs of one node behind one handle -- contiguous column
blocks per device, one engine and host thread per device, NCCL gather of the per-column balance terms at the end of a run.  The
library is loaded on first use only (it links libnccl); the single-device package never needs it."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from .bmi import LgarbError, _load

_glib = None


def group_library_path() -> Path:
    return Path(__file__).resolve().parent / "liblgarb_group.so"


def _gload():
    global _glib
    if _glib is not None:
        return _glib
    _load()  # liblgarb.so first (the group library depends on it)
    p = group_library_path()
    if not p.exists():
        raise LgarbError(f"{p} is missing: run `make -C lgar_c_b200/csrc` (needs nvcc and libnccl)")
    L = C.CDLL(str(p), mode=C.RTLD_GLOBAL)
    vp, i64, dp = C.c_void_p, C.c_int64, C.POINTER(C.c_double)
    sig = {
        "lgarb_group_create": ([C.POINTER(vp), C.c_int, C.POINTER(C.c_int)], C.c_int),
        "lgarb_group_destroy": ([vp], C.c_int),
        "lgarb_group_last_error": ([vp], C.c_char_p),
        "lgarb_group_initialize": ([vp, C.c_char_p, i64, vp, vp], C.c_int),
        "lgarb_group_finalize": ([vp], C.c_int),
        "lgarb_group_num_devices": ([vp], C.c_int),
        "lgarb_group_num_columns": ([vp], i64),
        "lgarb_group_shard": ([vp, C.c_int, C.POINTER(i64), C.POINTER(i64)], C.c_int),
        "lgarb_group_engine": ([vp, C.c_int], vp),
        "lgarb_group_update_steps_host": ([vp, i64, vp, vp, C.c_int, C.POINTER(C.c_char_p), C.POINTER(vp)], C.c_int),
        "lgarb_group_gather_mass_balance": ([vp, vp, dp, dp], C.c_int),
    }
    for name, (args, res) in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = res
    _glib = L
    return L


GROUP_SYMBOLS = ["lgarb_group_create", "lgarb_group_destroy", "lgarb_group_last_error", "lgarb_group_initialize", "lgarb_group_finalize",
                 "lgarb_group_num_devices", "lgarb_group_num_columns", "lgarb_group_shard", "lgarb_group_engine",
                 "lgarb_group_update_steps_host", "lgarb_group_gather_mass_balance"]


class LgarGroup:
    def __init__(self, devices):
        self._L = _gload()
        self._h = C.c_void_p()
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        if self._L.lgarb_group_create(C.byref(self._h), len(devices), devs) != 0:
            raise LgarbError(f"lgarb_group_create failed for devices {list(devices)}")
        self.n_columns = 0

    def _ck(self, rc):
        if rc != 0:
            raise LgarbError(self._L.lgarb_group_last_error(self._h).decode())

    def initialize(self, config_file, n_columns, layer_soil_type=None, layer_thickness_cm=None):
        st = None if layer_soil_type is None else np.ascontiguousarray(layer_soil_type, dtype=np.int32)
        tk = None if layer_thickness_cm is None else np.ascontiguousarray(layer_thickness_cm, dtype=np.float64)
        self._keep = (st, tk)
        self._ck(self._L.lgarb_group_initialize(self._h, str(config_file).encode(), int(n_columns), None if st is None else st.ctypes.data,
                                                None if tk is None else tk.ctypes.data))
        self.n_columns = int(n_columns)

    def shards(self):
        out = []
        for k in range(self._L.lgarb_group_num_devices(self._h)):
            a, b = C.c_int64(), C.c_int64()
            self._ck(self._L.lgarb_group_shard(self._h, k, C.byref(a), C.byref(b)))
            out.append((a.value, b.value))
        return out

    def update_steps_host(self, precip, pet, output_names=()):
        P = np.ascontiguousarray(precip, dtype=np.float64)
        E = np.ascontiguousarray(pet, dtype=np.float64)
        if P.ndim != 2 or P.shape[1] != self.n_columns or E.shape != P.shape:
            raise LgarbError(f"precip and pet must both be [n_steps][{self.n_columns}]")
        n = P.shape[0]
        outs = {nm: np.empty((n, self.n_columns)) for nm in output_names}
        names = (C.c_char_p * max(1, len(outs)))(*[nm.encode() for nm in outs])
        ptrs = (C.c_void_p * max(1, len(outs)))(*[a.ctypes.data for a in outs.values()])
        self._ck(self._L.lgarb_group_update_steps_host(self._h, n, P.ctypes.data, E.ctypes.data, len(outs), names, ptrs))
        return outs

    def gather_mass_balance(self):
        out = np.empty((self.n_columns, 16))
        tot = np.empty(16)
        ms = C.c_double()
        self._ck(self._L.lgarb_group_gather_mass_balance(self._h, out.ctypes.data, tot.ctypes.data_as(C.POINTER(C.c_double)), C.byref(ms)))
        return out, tot, ms.value

    def finalize(self):
        if self._h:
            self._L.lgarb_group_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.finalize()
        except Exception:
            pass
