"""The oracle against the unmodified reference executed live (oracle/_ref), including carried state
that the golden files do not hold.  CPU only; skipped when oracle/_ref is not built."""
import numpy as np
import pytest

from helpers import config_for, load_golden


@pytest.mark.parametrize("name,nsteps", [("Phillipsburg_with_reservoir", 1500), ("synth_0", 12), ("rand_3", 400), ("rand_5", 400)])
def test_oracle_tracks_reference_state(oracle, ref, name, nsteps, tmp_path):
    g = load_golden(name)
    cfg = config_for(name, tmp_path)
    p = oracle.params_from_config(cfg)
    s = oracle.new_state(p)
    h = ref.create(cfg)
    try:
        for t in range(nsteps):
            assert ref.update(h, g["precip"][t], g["pet"][t]) == 0
            st, out = oracle.update(p, s, g["precip"][t], g["pet"][t])
            assert st & 0xFFFF == 0
            rs = ref.state(h)
            mine = [s.volon_timestep_cm, s.precip_previous_timestep_cm, s.CR_fast_storage_cm, s.CR_slow_storage_cm, s.previous_AET,
                    s.previous_PET, s.previous_recharge, s.accumulated_PET, s.accumulated_free_drainage, s.timestep_h,
                    float(s.cache_fluxes), float(s.cache_count), float(s.runoff_in_prev_step)]
            assert list(rs[:13]) == mine, f"carried state differs at step {t}"
            assert rs[14] == s.timesteps
            assert np.array_equal(ref.scalars(h), out)
            fr, fl = ref.fronts(h)
            ofr, ofl = oracle.fronts_of(s)
            assert np.array_equal(fr, ofr) and np.array_equal(fl, ofl)
    finally:
        ref.destroy(h)


def test_ref_soil_table_matches_oracle(oracle, ref, tmp_path):
    from oracle.oracle_py import DATA
    cfg = config_for("Bushland", tmp_path)
    h = ref.create(cfg)
    table, n = oracle.read_soil_file(DATA / "vG_default_params.dat", 25)
    try:
        for i in range(1, n + 1):
            r = ref.soil(h, i)
            t = table[i]
            assert list(r) == [t.theta_r, t.theta_e, t.alpha, t.n, t.m, t.lam, t.psib, t.h_min, t.Ksat, t.theta_wp]
    finally:
        ref.destroy(h)
