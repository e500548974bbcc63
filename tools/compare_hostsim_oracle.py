#!/usr/bin/env python
"""Kernel text compiled for the host (tests/hostsim) vs the plain-C oracle, step by step."""
import sys, ctypes as C, tempfile
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle.oracle_py import Oracle, read_forcing, DATA, parse_config, resolve_data_path, absolutize_config, write_config

_dp = C.POINTER(C.c_double); _ip = C.POINTER(C.c_int)
def hostsim_lib():
    import subprocess
    d = Path(__file__).resolve().parent.parent/"tests"/"hostsim"
    subprocess.check_call(["g++","-std=c++17","-O2","-ffp-contract=off","-fPIC","-shared","-w","-o",str(d/"libhostsim.so"),str(d/"hostsim.cpp")])
    lib = C.CDLL(str(d/"libhostsim.so"))
    lib.hostsim_run_series.argtypes = [C.c_char_p,_ip,_dp,C.c_int,_dp,_dp,_dp,_ip,C.c_int,_dp,_ip,C.c_int,_ip,C.c_char_p,C.c_int]
    lib.hostsim_run_series.restype = C.c_int
    return lib

def run_hostsim(lib, cfg, P, E, cap=32, force=0):
    P = np.ascontiguousarray(P, dtype=np.float64); E = np.ascontiguousarray(E, dtype=np.float64)
    n = len(P)
    scal = np.zeros((n,12)); nf = np.zeros(n, dtype=np.int32); fr = np.zeros((n,cap,5)); fl = np.zeros((n,cap), dtype=np.int32)
    st = C.c_int(0); err = C.create_string_buffer(512)
    acfg = absolutize_config(cfg)
    done = lib.hostsim_run_series(str(acfg).encode(), None, None, n, P.ctypes.data_as(_dp), E.ctypes.data_as(_dp), scal.ctypes.data_as(_dp),
                                  nf.ctypes.data_as(_ip), cap, fr.ctypes.data_as(_dp), fl.ctypes.data_as(_ip), force, C.byref(st), err, 512)
    if done < 0: raise RuntimeError(err.value.decode())
    return dict(done=done, scalars=scal, nfronts=nf, fronts=fr, flags=fl, status=st.value)

def compare(lib, orc, cfg, P, E, label="", force=0):
    p = orc.params_from_config(cfg); s = orc.new_state(p)
    o = orc.run_series(p, s, P, E)
    h = run_hostsim(lib, cfg, P, E, force=force)
    n = min(o["done"], h["done"])
    ok_nf = np.array_equal(o["nfronts"][:n], h["nfronts"][:n])
    sd = np.abs(o["scalars"][:n] - h["scalars"][:n])
    biteq = np.array_equal(o["scalars"][:n], h["scalars"][:n]) and np.array_equal(o["fronts"][:n], h["fronts"][:n]) and np.array_equal(o["flags"][:n], h["flags"][:n])
    res = dict(label=label, done_o=o["done"], done_h=h["done"], status_h=h["status"], nf_equal=ok_nf, bit_equal=biteq, max_abs_scalar=float(sd.max()) if n else 0.0)
    if not biteq:
        bad = np.argwhere((o["scalars"][:n] != h["scalars"][:n]).any(axis=1) | (o["fronts"][:n] != h["fronts"][:n]).any(axis=(1,2)) | (o["flags"][:n] != h["flags"][:n]).any(axis=1))
        res["first_diff_step"] = int(bad[0][0])
    print(res)
    return res, o, h

if __name__ == "__main__":
    lib = hostsim_lib(); orc = Oracle()
    for nm in ["unittest.txt","config_lasam_synth_0.txt","config_lasam_synth_1.txt","config_lasam_synth_2.txt",
               "config_lasam_Phillipsburg.txt","config_lasam_Phillipsburg_with_reservoir.txt","config_lasam_Bushland.txt"]:
        cfg = DATA/"configs"/nm
        c = parse_config(cfg)
        if nm == "unittest.txt":
            P, E = np.array([1.896, 0.5, 0.0, 0.0, 3.0]), np.array([0.104, 0.1, 0.2, 0.3, 0.0])
        else:
            P, E = read_forcing(resolve_data_path(c["forcing_file"][0], cfg))
            end_h = float(c["endtime"][0]) * (1.0 if c["endtime"][1] in ("[h]","[hr]") else 1/3600.)
            n = int(end_h/(float(c["forcing_resolution"][0])/3600.)); P, E = P[:n], E[:n]
        compare(lib, orc, cfg, P, E, nm)
        compare(lib, orc, cfg, P, E, nm+" [force_all_active]", force=1)
