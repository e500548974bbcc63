"""oracle/oracle_py.py -- TEST INFRASTRUCTURE ONLY.

ctypes bindings for the two CPU checkers:

* ``Oracle``      -- oracle/liblgar_oracle.so, the plain-C restatement (oracle/lgar_oracle.c);
* ``RefHarness``  -- oracle/_ref/liblgar_ref.so, the UNMODIFIED reference compiled from
                     /root/reference by oracle/Makefile (travels to the GPU box as a built .so).

plus a small Python reader for the reference's ``key=value[unit]`` config files
(reference src/lgar.cxx:231-1016) used only to fill the oracle's parameter struct.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (lgar_c_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REPO = HERE.parent
DATA = REPO / "data"

MAX_FRONTS = 64
MAX_LAYERS = 8
MAX_GIUH = 64
NSCALARS = 12
SCALAR_NAMES = [
    "precipitation", "potential_evapotranspiration", "actual_evapotranspiration",
    "surface_runoff", "giuh_runoff", "soil_storage", "total_discharge", "infiltration",
    "percolation", "conceptual_reservoir_to_stream_discharge", "mass_balance",
    "conceptual_reservoir_storage",
]


class Soil(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("theta_r", "theta_e", "alpha", "n", "m", "lam", "psib", "h_min", "Ksat", "theta_wp")]


class Front(C.Structure):
    _fields_ = [("depth", C.c_double), ("theta", C.c_double), ("psi", C.c_double), ("K", C.c_double),
                ("dzdt", C.c_double), ("layer", C.c_int), ("to_bottom", C.c_int)]


class Params(C.Structure):
    _fields_ = [
        ("num_layers", C.c_int),
        ("cum_thickness", C.c_double * (MAX_LAYERS + 1)),
        ("soil", Soil * (MAX_LAYERS + 1)),
        ("initial_psi_cm", C.c_double),
        ("wilting_point_psi_cm", C.c_double), ("field_capacity_psi_cm", C.c_double),
        ("use_closed_form_G", C.c_int), ("adaptive_timestep", C.c_int), ("allow_flux_caching", C.c_int),
        ("free_drainage_enabled", C.c_int), ("free_drainage_to_CR", C.c_int), ("PET_affects_precip", C.c_int),
        ("invalid_soil_type", C.c_int),
        ("ponded_depth_max_cm", C.c_double),
        ("a", C.c_double), ("b", C.c_double), ("frac_to_CR", C.c_double), ("a_slow", C.c_double),
        ("b_slow", C.c_double), ("frac_slow", C.c_double), ("spf_factor", C.c_double),
        ("mbal_tol", C.c_double),
        ("timestep_h", C.c_double), ("forcing_resolution_h", C.c_double),
        ("nint", C.c_int), ("num_giuh", C.c_int),
        ("giuh", C.c_double * MAX_GIUH),
    ]


class State(C.Structure):
    _fields_ = [
        ("n", C.c_int),
        ("f", Front * MAX_FRONTS),
        ("volon_timestep_cm", C.c_double), ("precip_previous_timestep_cm", C.c_double),
        ("CR_fast_storage_cm", C.c_double), ("CR_slow_storage_cm", C.c_double),
        ("previous_AET", C.c_double), ("previous_PET", C.c_double), ("previous_recharge", C.c_double),
        ("accumulated_PET", C.c_double), ("accumulated_free_drainage", C.c_double),
        ("timestep_h", C.c_double),
        ("cache_fluxes", C.c_int), ("cache_count", C.c_int), ("runoff_in_prev_step", C.c_int),
        ("giuh_queue", C.c_double * (MAX_GIUH + 1)),
        ("local_mass_balance", C.c_double),
        ("volstart_cm", C.c_double), ("volprecip_cm", C.c_double), ("volin_cm", C.c_double),
        ("volend_cm", C.c_double), ("volon_cm", C.c_double), ("volrunoff_cm", C.c_double),
        ("volrunoff_giuh_cm", C.c_double), ("volrech_cm", C.c_double), ("volAET_cm", C.c_double),
        ("volPET_cm", C.c_double), ("volrunoff_CR_cm", C.c_double), ("volCRend_cm", C.c_double),
        ("volQ_cm", C.c_double), ("volQ_CR_cm", C.c_double),
        ("timesteps", C.c_longlong),
        ("status", C.c_int),
        ("n_theta_calls", C.c_longlong), ("n_tmb_iters", C.c_longlong), ("n_active_substeps", C.c_longlong),
        ("n_path", C.c_longlong * 24),
    ]


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def _as_dp(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _as_ip(a):
    return a.ctypes.data_as(_ip) if a is not None else None


def build_oracle(force: bool = False) -> Path:
    if os.environ.get("LGAR_ORACLE_SO"):  # an instrumented build of the same source (tests: the oracle under the sanitizers)
        return Path(os.environ["LGAR_ORACLE_SO"])
    so = HERE / "liblgar_oracle.so"
    src = HERE / "lgar_oracle.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["make", "-s", "-C", str(HERE), "oracle"])
    return so


def build_ref() -> Path | None:
    """(Re)build oracle/_ref from /root/reference when it is present; else keep the prebuilt one."""
    so = HERE / "_ref" / "liblgar_ref.so"
    if Path("/root/reference/src").is_dir():
        subprocess.check_call(["make", "-s", "-C", str(HERE), "ref"])
    return so if so.exists() else None


# ------------------------------------------------------------------------------------------------
# config file reader (mirrors InitFromConfigFile, reference src/lgar.cxx:231-1016)
# ------------------------------------------------------------------------------------------------
def _read_vector(v: str):
    return [float(x) for x in v.split(",")]


def parse_config(path: str | os.PathLike) -> dict:
    cfg: dict = {}
    for line in Path(path).read_text().splitlines():
        if "=" not in line:
            continue
        eq = line.find("=")
        key = line[:eq]
        lu = line.find("[")
        unit = line[lu:line.find("]") + 1] if lu >= 0 else ""
        val = line[eq + 1: lu] if lu >= 0 else line[eq + 1:]
        cfg[key] = (val, unit)
    return cfg


def _to_hours(val: float, unit: str) -> float:
    if unit in ("[s]", "[sec]", ""):
        return val / 3600
    if unit in ("[min]", "[minute]"):
        return val / 60
    if unit in ("[h]", "[hr]"):
        return val / 1.0
    return val


def _bool(cfg, key, default=False):
    if key not in cfg:
        return default
    v = cfg[key][0]
    return v in ("true", "1")


class Oracle:
    """The plain-C restatement, one column per (Params, State) pair."""

    def __init__(self):
        self.lib = C.CDLL(str(build_oracle()))
        L = self.lib
        L.lgo_soil_derive.argtypes = [C.POINTER(Soil)] + [C.c_double] * 6 + [C.c_int]
        L.lgo_read_soil_file.argtypes = [C.c_char_p, C.c_int, C.c_double, C.c_int, C.POINTER(Soil)]
        L.lgo_read_soil_file.restype = C.c_int
        L.lgo_init_state.argtypes = [C.POINTER(Params), C.POINTER(State)]
        L.lgo_update.argtypes = [C.POINTER(Params), C.POINTER(State), C.c_double, C.c_double, _dp]
        L.lgo_update.restype = C.c_int
        L.lgo_run_series.argtypes = [C.POINTER(Params), C.POINTER(State), C.c_int, _dp, _dp, _dp, _ip, C.c_int, _dp, _ip]
        L.lgo_run_series.restype = C.c_int
        L.lgo_global_balance.argtypes = [C.POINTER(State)]
        L.lgo_global_balance.restype = C.c_double
        L.lgo_theta_from_h.argtypes = [C.c_double, C.POINTER(Soil)]
        L.lgo_theta_from_h.restype = C.c_double
        for fn in ("lgo_Se_from_h", "lgo_h_from_Se"):
            getattr(L, fn).argtypes = [C.c_double] * 4
            getattr(L, fn).restype = C.c_double
        L.lgo_K_from_Se.argtypes = [C.c_double] * 3
        L.lgo_K_from_Se.restype = C.c_double
        L.lgo_Geff.argtypes = [C.c_int, C.c_double, C.c_double, C.POINTER(Soil), C.c_int]
        L.lgo_Geff.restype = C.c_double
        L.lgo_calc_mass_bal.argtypes = [_dp, C.POINTER(Front), C.c_int]
        L.lgo_calc_mass_bal.restype = C.c_double
        L.lgo_calc_aet.argtypes = [C.c_double] * 4 + [C.POINTER(Front), C.POINTER(Soil)]
        L.lgo_calc_aet.restype = C.c_double
        L.lgo_calc_CR_Q.argtypes = [C.c_double] * 7 + [_dp, _dp]
        L.lgo_calc_CR_Q.restype = C.c_double
        L.lgo_giuh.argtypes = [C.c_double, C.c_int, _dp, _dp]
        L.lgo_giuh.restype = C.c_double

    # -- soils ----------------------------------------------------------------------------------
    def read_soil_file(self, path, num_soil_types=25, wilting_point_psi_cm=15495.0, log_mode=False):
        table = (Soil * (num_soil_types + 1))()
        n = self.lib.lgo_read_soil_file(str(path).encode(), num_soil_types, wilting_point_psi_cm, int(log_mode), table)
        if n < 0:
            raise FileNotFoundError(path)
        return table, n

    def soil_from_values(self, theta_r, theta_e, alpha, n, Ksat, wilting_point_psi_cm=15495.0):
        s = Soil()
        self.lib.lgo_soil_derive(C.byref(s), theta_r, theta_e, alpha, n, Ksat, wilting_point_psi_cm, 0)
        return s

    # -- params ---------------------------------------------------------------------------------
    def params_from_config(self, config_path, soil_file=None, layer_soil_type=None, layer_thickness=None,
                           overrides=None) -> Params:
        cfg = parse_config(config_path)
        p = Params()
        thick = layer_thickness if layer_thickness is not None else _read_vector(cfg["layer_thickness"][0])
        p.num_layers = len(thick)
        cum = 0.0
        p.cum_thickness[0] = 0.0
        for i, t in enumerate(thick):
            cum = cum + t
            p.cum_thickness[i + 1] = cum
        p.initial_psi_cm = float(cfg["initial_psi"][0])
        p.wilting_point_psi_cm = float(cfg["wilting_point_psi"][0])
        p.field_capacity_psi_cm = float(cfg["field_capacity_psi"][0])
        p.use_closed_form_G = _bool(cfg, "use_closed_form_G")
        p.adaptive_timestep = _bool(cfg, "adaptive_timestep")
        p.allow_flux_caching = _bool(cfg, "allow_flux_caching")
        p.free_drainage_enabled = _bool(cfg, "free_drainage_enabled")
        p.free_drainage_to_CR = _bool(cfg, "free_drainage_to_CR")
        p.PET_affects_precip = _bool(cfg, "PET_affects_precip")
        log_mode = _bool(cfg, "log_mode")
        p.ponded_depth_max_cm = max(float(cfg["ponded_depth_max"][0]), 0.0) if "ponded_depth_max" in cfg else 0.0
        for k in ("a", "b", "a_slow", "b_slow", "frac_slow"):
            setattr(p, k, float(cfg[k][0]) if k in cfg else 0.0)
        p.frac_to_CR = float(cfg["frac_to_CR"][0]) if "frac_to_CR" in cfg else (
            float(cfg["frac_to_GW"][0]) if "frac_to_GW" in cfg else 0.0)
        p.spf_factor = float(cfg["spf_factor"][0]) if "spf_factor" in cfg else 0.98
        if log_mode:
            p.a = 10.0 ** p.a
            if "a_slow" in cfg:
                p.a_slow = 10.0 ** p.a_slow
        p.mbal_tol = float(cfg["mbal_tol"][0]) if "mbal_tol" in cfg else 10.0
        p.timestep_h = _to_hours(float(cfg["timestep"][0]), cfg["timestep"][1])
        p.forcing_resolution_h = _to_hours(float(cfg["forcing_resolution"][0]), cfg["forcing_resolution"][1])
        p.nint = 120
        ords = _read_vector(cfg["giuh_ordinates"][0])
        factor = int(1.0 / p.forcing_resolution_h)
        p.num_giuh = factor * len(ords)
        for i, o in enumerate(ords):
            for j in range(factor):
                p.giuh[j + i * factor] = o / float(factor)
        nst = min(int(cfg["max_valid_soil_types"][0]), 25) if "max_valid_soil_types" in cfg else 25
        sfile = soil_file if soil_file is not None else resolve_data_path(cfg["soil_params_file"][0], config_path)
        table, nread = self.read_soil_file(sfile, nst, p.wilting_point_psi_cm, log_mode)
        types = layer_soil_type if layer_soil_type is not None else [int(float(x)) for x in cfg["layer_soil_type"][0].split(",")]
        p.invalid_soil_type = 0
        for i, st in enumerate(types):
            if st > nst:
                p.invalid_soil_type = 1
                break
        if not p.invalid_soil_type:
            for i, st in enumerate(types):
                p.soil[i + 1] = table[st]
        if overrides:
            for k, v in overrides.items():
                setattr(p, k, v)
        return p

    def new_state(self, p: Params) -> State:
        s = State()
        self.lib.lgo_init_state(C.byref(p), C.byref(s))
        return s

    def update(self, p, s, precip, pet):
        out = np.zeros(NSCALARS)
        st = self.lib.lgo_update(C.byref(p), C.byref(s), float(precip), float(pet), _as_dp(out))
        return st, out

    def run_series(self, p, s, precip, pet, cap=32, want_fronts=True):
        precip = np.ascontiguousarray(precip, dtype=np.float64)
        pet = np.ascontiguousarray(pet, dtype=np.float64)
        n = len(precip)
        scal = np.zeros((n, NSCALARS))
        nf = np.zeros(n, dtype=np.int32)
        fr = np.zeros((n, cap, 5)) if want_fronts else None
        fl = np.zeros((n, cap), dtype=np.int32) if want_fronts else None
        done = self.lib.lgo_run_series(C.byref(p), C.byref(s), n, _as_dp(precip), _as_dp(pet), _as_dp(scal),
                                       _as_ip(nf), cap, _as_dp(fr), _as_ip(fl))
        return dict(done=done, scalars=scal, nfronts=nf, fronts=fr, flags=fl)

    @staticmethod
    def fronts_of(s: State):
        n = s.n
        arr = np.array([[s.f[i].depth, s.f[i].theta, s.f[i].psi, s.f[i].K, s.f[i].dzdt] for i in range(n)]).reshape(n, 5)
        flags = np.array([s.f[i].layer | (s.f[i].to_bottom << 8) for i in range(n)], dtype=np.int32)
        return arr, flags


def resolve_data_path(p: str, config_path) -> Path:
    """Config files point at ../data/x.dat or ./data/x.dat relative to where the reference is run;
    in this repo the same files live under data/."""
    name = Path(p).name
    for cand in (DATA / name, DATA / "forcing" / name, Path(config_path).parent / p, Path(p)):
        if cand.exists():
            return cand
    raise FileNotFoundError(p)


def absolutize_config(config_path, out_dir=None) -> Path:
    """Write a copy of a reference config whose soil_params_file / forcing_file are absolute paths
    into this repo's data/ directory (the reference opens them relative to its cwd)."""
    lines = []
    for line in Path(config_path).read_text().splitlines():
        if line.startswith("soil_params_file="):
            k, v = line.split("=", 1)
            line = f"{k}={resolve_data_path(v, config_path)}"
        if line.startswith("forcing_file="):  # only read by the standalone driver, never by Initialize
            k, v = line.split("=", 1)
            try:
                line = f"{k}={resolve_data_path(v, config_path)}"
            except FileNotFoundError:
                pass
        if line.startswith("verbosity="):
            line = "verbosity=none"
        lines.append(line)
    fd, tmp = tempfile.mkstemp(suffix=".txt", prefix="lgar_cfg_", dir=out_dir)
    with os.fdopen(fd, "w") as f:
        f.write("\n".join(lines) + "\n")
    return Path(tmp)


def write_config(base_config, out_path, **replace) -> Path:
    """Copy of a config with selected keys replaced (value strings include any [unit])."""
    seen = set()
    lines = []
    for line in Path(base_config).read_text().splitlines():
        if "=" in line:
            k = line.split("=", 1)[0]
            if k in replace:
                seen.add(k)
                if replace[k] is None:
                    continue
                line = f"{k}={replace[k]}"
        lines.append(line)
    for k, v in replace.items():
        if k not in seen and v is not None:
            lines.append(f"{k}={v}")
    Path(out_path).write_text("\n".join(lines) + "\n")
    return Path(out_path)


class RefHarness:
    """The unmodified reference behind oracle/ref_harness.cpp."""

    def __init__(self, so=None):
        if so is not None:
            so = Path(so)
        else:
            so = HERE / "_ref" / "liblgar_ref.so"
        if not so.exists():
            so = build_ref()
        if so is None or not Path(so).exists():
            raise FileNotFoundError("oracle/_ref/liblgar_ref.so (build with `make -C oracle ref`)")
        self.lib = C.CDLL(str(so))
        L = self.lib
        L.lgref_create.argtypes = [C.c_char_p]
        L.lgref_create.restype = C.c_void_p
        L.lgref_update.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.lgref_update.restype = C.c_int
        L.lgref_get_scalars.argtypes = [C.c_void_p, _dp]
        L.lgref_num_fronts.argtypes = [C.c_void_p]
        L.lgref_num_fronts.restype = C.c_int
        L.lgref_get_fronts.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp, _dp, _dp, _ip, _ip]
        L.lgref_get_fronts.restype = C.c_int
        L.lgref_get_bmi_fronts.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
        L.lgref_get_bmi_fronts.restype = C.c_int
        L.lgref_get_state.argtypes = [C.c_void_p, _dp]
        L.lgref_get_cumulative.argtypes = [C.c_void_p, _dp]
        L.lgref_get_soil.argtypes = [C.c_void_p, C.c_int, _dp]
        L.lgref_num_giuh.argtypes = [C.c_void_p]
        L.lgref_num_giuh.restype = C.c_int
        L.lgref_run_series.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp, _ip, C.c_int, _dp, _ip]
        L.lgref_run_series.restype = C.c_int
        L.lgref_time_columns.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, _dp, _dp, C.c_int, _ip, _dp, _dp]
        L.lgref_time_columns.restype = C.c_double
        L.lgref_destroy.argtypes = [C.c_void_p]
        if hasattr(L, "lgref_set_scalar"):
            L.lgref_set_scalar.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
            L.lgref_set_scalar.restype = C.c_int
            L.lgref_get_scalar.argtypes = [C.c_void_p, C.c_char_p, _dp]
            L.lgref_get_scalar.restype = C.c_int

    def set_scalar(self, h, name, value):
        """BmiLGAR::SetValue(name, &value) for a scalar double variable (calibratable parameters)."""
        if self.lib.lgref_set_scalar(h, name.encode(), float(value)) != 0:
            raise KeyError(name)

    def set_array(self, h, name, values):
        """BmiLGAR::SetValue(name, values) for an array variable (soil_temperature_profile)."""
        v = np.ascontiguousarray(values, dtype=np.float64)
        self.lib.lgref_set_array.argtypes = [C.c_void_p, C.c_char_p, _dp]
        if self.lib.lgref_set_array(h, name.encode(), _as_dp(v)) != 0:
            raise KeyError(name)

    def get_scalar(self, h, name):
        v = C.c_double(0.0)
        if self.lib.lgref_get_scalar(h, name.encode(), C.byref(v)) != 0:
            raise KeyError(name)
        return v.value

    def create(self, config_path):
        cfg = absolutize_config(config_path)
        try:
            h = self.lib.lgref_create(str(cfg).encode())
        finally:
            os.unlink(cfg)
        if not h:
            raise RuntimeError(f"reference failed to initialise from {config_path}")
        return h

    def destroy(self, h):
        self.lib.lgref_destroy(h)

    def update(self, h, precip, pet):
        return self.lib.lgref_update(h, float(precip), float(pet))

    def scalars(self, h):
        out = np.zeros(NSCALARS)
        self.lib.lgref_get_scalars(h, _as_dp(out))
        return out

    def fronts(self, h, cap=64):
        a = [np.zeros(cap) for _ in range(5)]
        layer = np.zeros(cap, dtype=np.int32)
        tb = np.zeros(cap, dtype=np.int32)
        n = self.lib.lgref_get_fronts(h, cap, *[_as_dp(x) for x in a], _as_ip(layer), _as_ip(tb))
        arr = np.stack(a, axis=1)[:n]
        return arr, (layer[:n] | (tb[:n] << 8)).astype(np.int32)

    def state(self, h):
        out = np.zeros(16)
        self.lib.lgref_get_state(h, _as_dp(out))
        return out

    def cumulative(self, h):
        out = np.zeros(16)
        self.lib.lgref_get_cumulative(h, _as_dp(out))
        return out

    def soil(self, h, idx):
        out = np.zeros(10)
        self.lib.lgref_get_soil(h, idx, _as_dp(out))
        return out

    def run_series(self, h, precip, pet, cap=32, want_fronts=True):
        precip = np.ascontiguousarray(precip, dtype=np.float64)
        pet = np.ascontiguousarray(pet, dtype=np.float64)
        n = len(precip)
        scal = np.zeros((n, NSCALARS))
        nf = np.zeros(n, dtype=np.int32)
        fr = np.zeros((n, cap, 5)) if want_fronts else None
        fl = np.zeros((n, cap), dtype=np.int32) if want_fronts else None
        done = self.lib.lgref_run_series(h, n, _as_dp(precip), _as_dp(pet), _as_dp(scal), _as_ip(nf), cap,
                                         _as_dp(fr), _as_ip(fl))
        return dict(done=done, scalars=scal, nfronts=nf, fronts=fr, flags=fl)


def read_forcing(path):
    """Forcing CSV: header line, then ``time,P(mm/h),PET(mm/h)`` (reference src/bmi_main_lgar.cxx:183-262)."""
    P, E = [], []
    with open(path) as f:
        f.readline()
        for line in f:
            parts = line.strip().split(",")
            if len(parts) >= 3:
                P.append(float(parts[1]))
                E.append(float(parts[2]))
    return np.array(P), np.array(E)
