"""ctypes mirror of include/lgarb.h: the reference's BMI verbs over a column batch.

Method names follow the reference class (``Initialize/Update/UpdateUntil/GetValue/SetValue/Finalize``,
reference src/bmi_lgar.cxx) in BMI-Python snake case; arrays are numpy (host) or raw device pointers
(``*_device``, e.g. ``torch.Tensor.data_ptr()``).  Errors raise ``LgarbError`` with the engine's
message instead of the reference's ``std::runtime_error`` / ``abort()``.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
MAX_FRONTS = 32
MAX_LAYERS = 4
STATUS_FAIL_MASK = 0xFFFF
OUTPUT_SCALARS = [
    "precipitation", "potential_evapotranspiration", "actual_evapotranspiration", "surface_runoff",
    "giuh_runoff", "soil_storage", "total_discharge", "infiltration", "percolation",
    "conceptual_reservoir_to_stream_discharge", "mass_balance", "conceptual_reservoir_storage",
]


def _config_num_layers(config_file):
    """number of entries of layer_thickness in a reference configuration file (reference src/lgar.cxx:321-349)"""
    try:
        for line in open(config_file):
            if line.split("=")[0].strip() == "layer_thickness":
                return len([t for t in line.split("=", 1)[1].split("[")[0].split(",") if t.strip()])
    except OSError:
        pass
    return -1


# lgar_c_b200/csrc/lgar_types.h LGAR_PATH_* (= oracle/lgar_oracle.h LGO_PATH_*)
PATH_NAMES = ("merge", "cross_layer", "cross_bottom", "dry_over_wet", "theta_r_delete", "saturated_depth", "cross_layer_bottom", "surficial",
              "clean_redundant", "search", "search_converged", "search_step_exit", "search_nochange", "search_maxiter", "search_psi_upper",
              "search_psi_zero", "corr_search", "cached_step", "active_step")


class LgarbError(RuntimeError):
    pass


def library_path() -> Path:
    import os
    alt = os.environ.get("LGARB_LIBRARY")  # experiments with differently compiled builds of the same sources
    return Path(alt) if alt else HERE / "liblgarb.so"


def build_library(force: bool = False) -> Path:
    """Compile the CUDA extension in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    so = library_path()
    srcs = list((HERE / "csrc").glob("*")) + [HERE.parent / "include" / "lgarb.h"]
    newest = max(p.stat().st_mtime for p in srcs)
    if force or not so.exists() or so.stat().st_mtime < newest:
        subprocess.check_call(["make", "-s", "-C", str(HERE / "csrc")])
    return so


_lib = None


def _load():
    global _lib
    if _lib is not None:
        return _lib
    so = library_path()
    if not so.exists():
        raise LgarbError(f"{so} is missing: build it with lgar_c_b200.build_library() / make -C lgar_c_b200/csrc "
                         "(there is no CPU fallback)")
    L = C.CDLL(str(so))
    vp, cp, i64, dp = C.c_void_p, C.c_char_p, C.c_int64, C.POINTER(C.c_double)
    sig = {
        "lgarb_create": ([C.POINTER(vp)], C.c_int),
        "lgarb_destroy": ([vp], C.c_int),
        "lgarb_initialize": ([vp, cp, i64, C.c_int], C.c_int),
        "lgarb_initialize_columns": ([vp, cp, i64, C.c_int, vp, vp], C.c_int),
        "lgarb_finalize": ([vp], C.c_int),
        "lgarb_last_error": ([vp], cp),
        "lgarb_update": ([vp], C.c_int),
        "lgarb_update_until": ([vp, C.c_double], C.c_int),
        "lgarb_update_steps_device": ([vp, i64, vp, vp, i64, vp, i64, vp], C.c_int),
        "lgarb_update_steps_host": ([vp, i64, vp, vp, C.c_int, vp, vp], C.c_int),
        "lgarb_update_steps_host_pitched": ([vp, i64, vp, vp, i64, C.c_int, vp, vp, i64], C.c_int),
        "lgarb_get_global_mass_balance_device": ([vp, vp], C.c_int),
        "lgarb_update_steps_host_stream": ([vp, i64, i64, C.c_int, vp, vp, C.c_int, vp, C.c_int, vp], C.c_int),
        "lgarb_synchronize": ([vp], C.c_int),
        "lgarb_set_value": ([vp, cp, vp], C.c_int),
        "lgarb_set_value_range": ([vp, cp, i64, i64, vp], C.c_int),
        "lgarb_set_value_device": ([vp, cp, vp], C.c_int),
        "lgarb_set_value_at_indices": ([vp, cp, vp, C.c_int, vp], C.c_int),
        "lgarb_get_value": ([vp, cp, vp], C.c_int),
        "lgarb_get_value_range": ([vp, cp, i64, i64, vp], C.c_int),
        "lgarb_get_value_device": ([vp, cp, vp], C.c_int),
        "lgarb_get_value_ptr": ([vp, cp, C.POINTER(vp)], C.c_int),
        "lgarb_get_value_at_indices": ([vp, cp, vp, vp, C.c_int], C.c_int),
        "lgarb_get_component_name": ([vp, cp, C.c_int], C.c_int),
        "lgarb_get_input_item_count": ([vp, C.POINTER(C.c_int)], C.c_int),
        "lgarb_get_output_item_count": ([vp, C.POINTER(C.c_int)], C.c_int),
        "lgarb_get_input_var_name": ([vp, C.c_int], cp),
        "lgarb_get_output_var_name": ([vp, C.c_int], cp),
        "lgarb_get_var_grid": ([vp, cp, C.POINTER(C.c_int)], C.c_int),
        "lgarb_get_var_type": ([vp, cp, cp, C.c_int], C.c_int),
        "lgarb_get_var_itemsize": ([vp, cp, C.POINTER(C.c_int)], C.c_int),
        "lgarb_get_var_units": ([vp, cp, cp, C.c_int], C.c_int),
        "lgarb_get_var_nbytes": ([vp, cp, C.POINTER(i64)], C.c_int),
        "lgarb_get_var_location": ([vp, cp, cp, C.c_int], C.c_int),
        "lgarb_get_current_time": ([vp, dp], C.c_int),
        "lgarb_get_start_time": ([vp, dp], C.c_int),
        "lgarb_get_end_time": ([vp, dp], C.c_int),
        "lgarb_get_time_step": ([vp, dp], C.c_int),
        "lgarb_get_time_units": ([vp, cp, C.c_int], C.c_int),
        "lgarb_get_grid_rank": ([vp, C.c_int, C.POINTER(C.c_int)], C.c_int),
        "lgarb_get_grid_size": ([vp, C.c_int, C.POINTER(i64)], C.c_int),
        "lgarb_get_grid_type": ([vp, C.c_int, cp, C.c_int], C.c_int),
        "lgarb_get_num_columns": ([vp], i64),
        "lgarb_get_num_layers": ([vp], C.c_int),
        "lgarb_get_num_giuh_ordinates": ([vp], C.c_int),
        "lgarb_output_scalar_name": ([C.c_int], cp),
        "lgarb_get_status": ([vp, i64, i64, vp], C.c_int),
        "lgarb_get_fronts": ([vp, i64, i64] + [vp] * 8, C.c_int),
        "lgarb_get_global_mass_balance": ([vp, i64, i64, vp], C.c_int),
        "lgarb_get_soil_row": ([vp, C.c_int, dp], C.c_int),
        "lgarb_set_force_all_active": ([vp, C.c_int], C.c_int),
        "lgarb_set_window_mode": ([vp, C.c_int], C.c_int),
        "lgarb_set_wave_params": ([vp, C.c_int, C.c_int, C.c_int], C.c_int),
        "lgarb_set_work_trace": ([vp, C.c_int], C.c_int),
        "lgarb_get_work_trace": ([vp, i64, i64, vp, vp], C.c_int),
        "lgarb_get_stage_profile": ([vp, vp], C.c_int),
        "lgarb_get_warp_times": ([vp, vp], C.c_int),
        "lgarb_get_launch_count": ([vp], i64),
        "lgarb_get_last_active_count": ([vp], i64),
        "lgarb_get_stream": ([vp], vp),
        "lgarb_measure_fp64_peak": ([vp, C.c_double, dp], C.c_int),
        "lgarb_debug_pow": ([vp, C.c_int64, vp, vp, vp, C.c_int], C.c_int),
        "lgarb_debug_functions": ([vp, C.c_int, C.c_int, C.c_int64, vp, vp, vp], C.c_int),
        "lgarb_set_path_counters": ([vp, C.c_int], C.c_int),
        "lgarb_set_stage_timing": ([vp, C.c_int], C.c_int),
        "lgarb_get_stage_times": ([vp, C.POINTER(C.c_double), C.POINTER(C.c_int64)], C.c_int),
        "lgarb_get_path_counters": ([vp, C.POINTER(C.c_uint64), C.c_int], C.c_int),
        "lgarb_set_profiling": ([vp, C.c_int], C.c_int),
        "lgarb_get_profile": ([vp, dp], C.c_int),
    }
    for name, (args, res) in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = res
    _lib = L
    return L


C_ABI_SYMBOLS = None  # filled lazily by tests from include/lgarb.h


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class BmiLGARBatch:
    """Batched counterpart of the reference's ``BmiLGAR`` (one object = n_columns soil columns)."""

    def __init__(self):
        self._L = _load()
        h = C.c_void_p()
        if self._L.lgarb_create(C.byref(h)) != 0:
            raise LgarbError("lgarb_create failed")
        self._h = h
        self.n_columns = 0
        self.num_layers = 0

    # -- plumbing -------------------------------------------------------------------------------
    def _ck(self, rc):
        if rc != 0:
            raise LgarbError(self._L.lgarb_last_error(self._h).decode())

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._L.lgarb_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # -- life cycle (reference Initialize / Finalize, src/bmi_lgar.cxx:72,1080) -------------------
    def initialize(self, config_file, n_columns=1, device=0, layer_soil_type=None, layer_thickness_cm=None):
        st = tk = None
        if layer_soil_type is not None:
            st = np.ascontiguousarray(layer_soil_type, dtype=np.int32)
            if st.shape[0] != n_columns:
                raise LgarbError("layer_soil_type must be [n_columns][num_layers]")
        if layer_thickness_cm is not None:
            tk = np.ascontiguousarray(layer_thickness_cm, dtype=np.float64)
            if tk.shape[0] != n_columns:
                raise LgarbError("layer_thickness_cm must be [n_columns][num_layers]")
        if st is not None and tk is not None and (st.ndim != 2 or st.shape != tk.shape):
            raise LgarbError("layer_soil_type and layer_thickness_cm must both be [n_columns][num_layers]")
        for a in (st, tk):
            # the C side indexes [column * num_layers + layer] with num_layers taken from the configuration file: check it
            # BEFORE the call so that a wrong second dimension is an error and not an out-of-bounds read
            if a is not None and (a.ndim != 2 or a.shape[1] != (_config_num_layers(config_file) if _config_num_layers(config_file) > 0 else a.shape[1])):
                raise LgarbError("per-column layer arrays must be [n_columns][num_layers of the configuration]")
        self._ck(self._L.lgarb_initialize_columns(self._h, str(config_file).encode(), int(n_columns), int(device), _ptr(st), _ptr(tk)))
        self.n_columns = int(n_columns)
        self.num_layers = self._L.lgarb_get_num_layers(self._h)

    def finalize(self):
        self._ck(self._L.lgarb_finalize(self._h))

    # -- time stepping (reference Update / UpdateUntil, src/bmi_lgar.cxx:133,898) ------------------
    def update(self):
        self._ck(self._L.lgarb_update(self._h))

    def update_until(self, t):
        self._ck(self._L.lgarb_update_until(self._h, float(t)))

    def update_steps_device(self, n_steps, d_precip, d_pet, series_stride, sinks=None, sink_stride=0, nfronts_sink=None):
        """d_precip/d_pet: device pointers (int) of [n_steps][series_stride] f64 series (mm/h).
        sinks: optional dict {output name: device pointer of [n_steps][sink_stride]};
        nfronts_sink: optional device pointer of an int32 [n_steps][sink_stride] buffer."""
        arr = None
        if sinks:
            arr = (C.c_void_p * len(OUTPUT_SCALARS))()
            for k, name in enumerate(OUTPUT_SCALARS):
                arr[k] = sinks.get(name, None)
        self._ck(self._L.lgarb_update_steps_device(self._h, int(n_steps), C.c_void_p(int(d_precip)), C.c_void_p(int(d_pet)),
                                                    int(series_stride), arr, int(sink_stride),
                                                    C.c_void_p(int(nfronts_sink)) if nfronts_sink else None))

    def update_steps_host(self, precip, pet, outputs=()):
        """precip/pet: host arrays [n_steps][n_columns] (mm/h).  outputs: {name: host array [n_steps][n_columns]}
        filled in place (or a list of names: arrays are allocated and returned)."""
        P = np.ascontiguousarray(precip, dtype=np.float64)
        E = np.ascontiguousarray(pet, dtype=np.float64)
        self._check_series(P, E)
        n_steps = P.shape[0]
        if not isinstance(outputs, dict):
            outputs = {n: np.empty((n_steps, self.n_columns)) for n in outputs}
        names = list(outputs)
        self._check_outputs(outputs, n_steps, np.float64)
        cn = (C.c_char_p * len(names))(*[n.encode() for n in names])
        cp = (C.c_void_p * len(names))(*[outputs[n].ctypes.data for n in names])
        self._ck(self._L.lgarb_update_steps_host(self._h, int(n_steps), _ptr(P), _ptr(E), len(names), cn, cp))
        return outputs

    def update_steps_host_block(self, precip, pet, col0, output_names=()):
        """This engine as the owner of the columns [col0, col0 + n_columns) of GLOBAL host arrays [n_steps][n_total]
        (lgarb_update_steps_host_pitched): forcing read from, outputs written into that block of columns.  Returns
        {name: global array} with only the block filled (NaN elsewhere)."""
        P = np.ascontiguousarray(precip, dtype=np.float64)
        E = np.ascontiguousarray(pet, dtype=np.float64)
        if P.ndim != 2 or E.shape != P.shape or col0 < 0 or col0 + self.n_columns > P.shape[1]:
            raise LgarbError("precip and pet must both be [n_steps][n_total] with the block inside")
        n_steps, ntot = P.shape
        outs = {n: np.full((n_steps, ntot), np.nan) for n in output_names}
        cn = (C.c_char_p * max(1, len(outs)))(*[n.encode() for n in outs])
        cp = (C.c_void_p * max(1, len(outs)))(*[a.ctypes.data + 8 * col0 for a in outs.values()])
        self._ck(self._L.lgarb_update_steps_host_pitched(self._h, int(n_steps), P.ctypes.data + 8 * col0, E.ctypes.data + 8 * col0, int(ntot), len(outs), cn, cp, int(ntot)))
        return outs

    def update_steps_host_stream(self, precip, pet, outputs=(), chunk_steps=0, output_dtype=np.float64):
        """Streamed variant of update_steps_host (lgarb_update_steps_host_stream): the series are cut into chunks of
        `chunk_steps` hours and the copies of neighbouring chunks overlap the computation.  precip/pet may be float64
        or float32 host arrays [n_steps][n_columns] (both the same type); outputs as in update_steps_host, of dtype
        `output_dtype` (float64 or float32)."""
        P = np.ascontiguousarray(precip)
        in32 = P.dtype == np.float32
        P = np.ascontiguousarray(P, dtype=np.float32 if in32 else np.float64)
        E = np.ascontiguousarray(pet, dtype=P.dtype)
        self._check_series(P, E)
        n_steps = P.shape[0]
        out32 = np.dtype(output_dtype) == np.float32
        if not isinstance(outputs, dict):
            outputs = {n: np.empty((n_steps, self.n_columns), dtype=np.float32 if out32 else np.float64) for n in outputs}
        names = list(outputs)
        self._check_outputs(outputs, n_steps, np.float32 if out32 else np.float64)
        cn = (C.c_char_p * len(names))(*[n.encode() for n in names])
        cp = (C.c_void_p * len(names))(*[outputs[n].ctypes.data for n in names])
        self._ck(self._L.lgarb_update_steps_host_stream(self._h, int(n_steps), int(chunk_steps), 1 if in32 else 0, _ptr(P), _ptr(E),
                                                        len(names), cn, 1 if out32 else 0, cp))
        return outputs

    def _check_series(self, P, E):
        # a wrong shape would be an out-of-bounds read on the native side
        if P.ndim != 2 or P.shape[1] != self.n_columns or E.shape != P.shape:
            raise LgarbError(f"precip and pet must both be [n_steps][{self.n_columns}], got {P.shape} and {E.shape}")

    def _check_outputs(self, outputs, n_steps, dtype):
        for n, a in outputs.items():
            if not isinstance(a, np.ndarray) or a.dtype != dtype or not a.flags.c_contiguous or a.shape != (n_steps, self.n_columns):
                raise LgarbError(f"output array {n}: must be a C-contiguous {np.dtype(dtype).name} array of shape ({n_steps}, {self.n_columns})")

    def synchronize(self):
        self._ck(self._L.lgarb_synchronize(self._h))

    # -- variables (reference GetValue / SetValue, src/bmi_lgar.cxx:1307,1455) ---------------------
    def set_value(self, name, src):
        a = np.ascontiguousarray(src, dtype=np.float64)
        if name == "soil_temperature_profile":     # [n_columns][num_cells_temp] (grid 4)
            want = self.get_grid_size(4)
        else:
            want = self.n_columns
        if a.size != want:
            raise LgarbError(f"{name}: expected {want} items, got {a.size}")
        self._ck(self._L.lgarb_set_value(self._h, name.encode(), _ptr(a)))

    def set_value_device(self, name, d_ptr):
        self._ck(self._L.lgarb_set_value_device(self._h, name.encode(), C.c_void_p(int(d_ptr))))

    def set_value_at_indices(self, name, inds, src):
        i = np.ascontiguousarray(inds, dtype=np.int32)
        a = np.ascontiguousarray(src, dtype=np.float64)
        self._ck(self._L.lgarb_set_value_at_indices(self._h, name.encode(), _ptr(i), len(i), _ptr(a)))

    def _alloc_for(self, name, ncol):
        if name == "soil_num_wetting_fronts":
            return np.empty(ncol, dtype=np.int32)
        if name in ("soil_moisture_wetting_fronts", "soil_depth_wetting_fronts"):
            return np.empty((ncol, MAX_FRONTS), dtype=np.float64)
        if name == "soil_depth_layers":
            return np.empty((ncol, self.num_layers), dtype=np.float64)
        return np.empty(ncol, dtype=np.float64)

    def get_value(self, name, dest=None):
        if dest is None:
            dest = self._alloc_for(name, self.n_columns)
        self._ck(self._L.lgarb_get_value(self._h, name.encode(), _ptr(dest)))
        return dest

    def get_value_range(self, name, col0, ncol):
        dest = self._alloc_for(name, ncol)
        self._ck(self._L.lgarb_get_value_range(self._h, name.encode(), int(col0), int(ncol), _ptr(dest)))
        return dest

    def get_value_device(self, name, d_ptr):
        self._ck(self._L.lgarb_get_value_device(self._h, name.encode(), C.c_void_p(int(d_ptr))))

    def get_value_ptr(self, name) -> int:
        p = C.c_void_p()
        self._ck(self._L.lgarb_get_value_ptr(self._h, name.encode(), C.byref(p)))
        return int(p.value)

    def get_value_at_indices(self, name, inds):
        i = np.ascontiguousarray(inds, dtype=np.int32)
        dest = np.empty(len(i), dtype=np.int32 if name == "soil_num_wetting_fronts" else np.float64)
        self._ck(self._L.lgarb_get_value_at_indices(self._h, name.encode(), _ptr(dest), _ptr(i), len(i)))
        return dest

    # -- metadata ---------------------------------------------------------------------------------
    def _str(self, fn, *args):
        buf = C.create_string_buffer(256)
        self._ck(fn(self._h, *args, buf, 256))
        return buf.value.decode()

    def get_component_name(self):
        return self._str(self._L.lgarb_get_component_name)

    def get_input_item_count(self):
        n = C.c_int()
        self._ck(self._L.lgarb_get_input_item_count(self._h, C.byref(n)))
        return n.value

    def get_output_item_count(self):
        n = C.c_int()
        self._ck(self._L.lgarb_get_output_item_count(self._h, C.byref(n)))
        return n.value

    def get_input_var_names(self):
        return [self._L.lgarb_get_input_var_name(self._h, i).decode() for i in range(self.get_input_item_count())]

    def get_output_var_names(self):
        return [self._L.lgarb_get_output_var_name(self._h, i).decode() for i in range(self.get_output_item_count())]

    def get_var_grid(self, name):
        g = C.c_int()
        self._ck(self._L.lgarb_get_var_grid(self._h, name.encode(), C.byref(g)))
        return g.value

    def get_var_type(self, name):
        return self._str(self._L.lgarb_get_var_type, name.encode())

    def get_var_units(self, name):
        return self._str(self._L.lgarb_get_var_units, name.encode())

    def get_var_location(self, name):
        return self._str(self._L.lgarb_get_var_location, name.encode())

    def get_var_itemsize(self, name):
        n = C.c_int()
        self._ck(self._L.lgarb_get_var_itemsize(self._h, name.encode(), C.byref(n)))
        return n.value

    def get_var_nbytes(self, name):
        n = C.c_int64()
        self._ck(self._L.lgarb_get_var_nbytes(self._h, name.encode(), C.byref(n)))
        return n.value

    def _dbl(self, fn):
        v = C.c_double()
        self._ck(fn(self._h, C.byref(v)))
        return v.value

    def get_current_time(self):
        return self._dbl(self._L.lgarb_get_current_time)

    def get_start_time(self):
        return self._dbl(self._L.lgarb_get_start_time)

    def get_end_time(self):
        return self._dbl(self._L.lgarb_get_end_time)

    def get_time_step(self):
        return self._dbl(self._L.lgarb_get_time_step)

    def get_time_units(self):
        return self._str(self._L.lgarb_get_time_units)

    def get_grid_rank(self, grid):
        n = C.c_int()
        self._ck(self._L.lgarb_get_grid_rank(self._h, int(grid), C.byref(n)))
        return n.value

    def get_grid_size(self, grid):
        n = C.c_int64()
        self._ck(self._L.lgarb_get_grid_size(self._h, int(grid), C.byref(n)))
        return n.value

    def get_grid_type(self, grid):
        return self._str(self._L.lgarb_get_grid_type, int(grid))

    # -- batch-only services ------------------------------------------------------------------------
    def get_status(self, col0=0, ncol=None):
        ncol = self.n_columns - col0 if ncol is None else ncol
        out = np.empty(ncol, dtype=np.int32)
        self._ck(self._L.lgarb_get_status(self._h, int(col0), int(ncol), _ptr(out)))
        return out

    def get_fronts(self, col0=0, ncol=None):
        """Raw front state: dict of [ncol][MAX_FRONTS] arrays (cm, cm/h) + nfronts[ncol]."""
        ncol = self.n_columns - col0 if ncol is None else ncol
        d = {k: np.empty((ncol, MAX_FRONTS)) for k in ("depth", "theta", "psi", "K", "dzdt")}
        layer = np.empty((ncol, MAX_FRONTS), dtype=np.int32)
        tb = np.empty((ncol, MAX_FRONTS), dtype=np.int32)
        nf = np.empty(ncol, dtype=np.int32)
        self._ck(self._L.lgarb_get_fronts(self._h, int(col0), int(ncol), _ptr(d["depth"]), _ptr(d["theta"]), _ptr(d["psi"]),
                                          _ptr(d["K"]), _ptr(d["dzdt"]), _ptr(layer), _ptr(tb), _ptr(nf)))
        d.update(layer=layer, to_bottom=tb, nfronts=nf)
        return d

    def get_global_mass_balance(self, col0=0, ncol=None):
        ncol = self.n_columns - col0 if ncol is None else ncol
        out = np.empty((ncol, 16))
        self._ck(self._L.lgarb_get_global_mass_balance(self._h, int(col0), int(ncol), _ptr(out)))
        return out

    def get_soil_row(self, soil):
        out = np.empty(10)
        self._ck(self._L.lgarb_get_soil_row(self._h, int(soil), out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def set_force_all_active(self, on):
        self._ck(self._L.lgarb_set_force_all_active(self._h, int(bool(on))))

    def set_window_mode(self, mode):
        """Scheduler of the multi-step verbs: 3 six stage kernels with queues between them (default); 1 warp-tile window
        kernel; 0 hour-by-hour kernel pairs.  All bit-identical.  (2, 4, 5 of earlier rounds were measured slower and removed.)"""
        self._ck(self._L.lgarb_set_window_mode(self._h, int(mode)))

    def set_wave_params(self, search_quantum=0, front_search_pairs=0, finish_below=-1):
        self._ck(self._L.lgarb_set_wave_params(self._h, int(search_quantum), int(front_search_pairs), int(finish_below)))

    def set_work_trace(self, on):
        self._ck(self._L.lgarb_set_work_trace(self._h, int(bool(on))))

    def get_work_trace(self, col0=0, ncol=None):
        ncol = self.n_columns - col0 if ncol is None else ncol
        a = np.empty(ncol, dtype=np.uint32)
        b = np.empty(ncol, dtype=np.uint32)
        self._ck(self._L.lgarb_get_work_trace(self._h, int(col0), int(ncol), _ptr(a), _ptr(b)))
        return a, b

    def get_stage_profile(self):
        a = np.zeros(15, dtype=np.uint64)
        self._ck(self._L.lgarb_get_stage_profile(self._h, _ptr(a)))
        names = ["FETCH", "HEAD", "FRONT", "SEARCH", "TAIL"]
        return {n: dict(cycles=int(a[k]), turns=int(a[5 + k]), lanes=int(a[10 + k])) for k, n in enumerate(names)}

    def get_warp_times(self):
        a = np.zeros(16384, dtype=np.uint32)
        self._ck(self._L.lgarb_get_warp_times(self._h, _ptr(a)))
        return a[:8192], a[8192:]

    def get_launch_count(self):
        return self._L.lgarb_get_launch_count(self._h)

    def get_last_active_count(self):
        return self._L.lgarb_get_last_active_count(self._h)

    def set_profiling(self, on):
        self._ck(self._L.lgarb_set_profiling(self._h, int(bool(on))))

    def get_profile(self):
        """dict(ms_stream, ms_active, steps, active_colsteps, front_sum) accumulated since set_profiling(True)."""
        out = np.zeros(6)
        self._ck(self._L.lgarb_get_profile(self._h, out.ctypes.data_as(C.POINTER(C.c_double))))
        return dict(ms_stream=out[0], ms_active=out[1], steps=int(out[2]), active_colsteps=int(out[3]), front_sum=int(out[4]),
                    ms_window=out[5])

    def measure_fp64_peak(self, ms: float = 50.0) -> float:
        """DFMA microbenchmark on the engine's device: TFLOP/s of the FP64 pipe (secondary roofline)."""
        out = C.c_double(0.0)
        self._ck(self._L.lgarb_measure_fp64_peak(self._h, float(ms), C.byref(out)))
        return out.value

    def set_stage_timing(self, on=True):
        self._ck(self._L.lgarb_set_stage_timing(self._h, int(bool(on))))

    def get_stage_times(self):
        """{stage: (milliseconds, launches)} accumulated since set_stage_timing(True): FETCH, HEAD, FRONT, SEARCH, TAIL, CORR, FINISH"""
        ms = (C.c_double * 7)()
        n = (C.c_int64 * 7)()
        self._ck(self._L.lgarb_get_stage_times(self._h, ms, n))
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(("FETCH", "HEAD", "FRONT", "SEARCH", "TAIL", "CORR", "FINISH"))}

    def set_path_counters(self, on=True):
        self._ck(self._L.lgarb_set_path_counters(self._h, int(bool(on))))

    def get_path_counters(self):
        """dict name -> count of the rare paths of the reference taken since set_path_counters(True) (lgar_types.h LGAR_PATH_*)"""
        out = (C.c_uint64 * len(PATH_NAMES))()
        self._ck(self._L.lgarb_get_path_counters(self._h, out, len(PATH_NAMES)))
        return dict(zip(PATH_NAMES, [int(x) for x in out]))

    def debug_pow(self, x, y, which=0):
        """The engine's device pow on host arrays (lgarb_debug_pow): for the test that compares it with the host's pow."""
        x = np.ascontiguousarray(x, dtype=np.float64).ravel()
        y = np.ascontiguousarray(y, dtype=np.float64).ravel()
        if x.shape != y.shape:
            raise LgarbError("x and y must have the same number of items")
        out = np.empty_like(x)
        self._ck(self._L.lgarb_debug_pow(self._h, x.size, _ptr(x), _ptr(y), _ptr(out), int(which)))
        return out

    def debug_functions(self, fn, a, b=None, layer=1):
        """Leaf functions of the path on the engine's current state, one item per column (lgarb_debug_functions)."""
        a = np.ascontiguousarray(a, dtype=np.float64).ravel()
        b = np.zeros_like(a) if b is None else np.ascontiguousarray(b, dtype=np.float64).ravel()
        if a.shape != b.shape or a.size > self.n_columns:
            raise LgarbError("a and b must have the same number of items, at most n_columns")
        out = np.empty_like(a)
        self._ck(self._L.lgarb_debug_functions(self._h, int(fn), int(layer), a.size, _ptr(a), _ptr(b), _ptr(out)))
        return out

    def get_stream(self) -> int:
        return int(self._L.lgarb_get_stream(self._h) or 0)
