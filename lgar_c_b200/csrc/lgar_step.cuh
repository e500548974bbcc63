// lgar_c_b200/csrc/lgar_step.cuh -- per-column bodies of the two step kernels: HBM <-> thread state
// marshalling (structure-of-arrays, column fastest), the decision which kernel finishes a column,
// and the per-step epilogue (GIUH, outputs, cumulative volumes).  Kept apart from lgar_kernels.cu so
// that tests/hostsim can run exactly this text on a CPU against the oracle (test infrastructure).
#ifndef LGAR_STEP_CUH
#define LGAR_STEP_CUH

#include "lgar_kernels.h"
#include "lgar_physics.cuh"

#ifdef LGAR_HOSTSIM
#define LG_NAN (std::nan(""))
#else
#include <math_constants.h>
#define LG_NAN CUDART_NAN
#endif

LG_FN_NOINLINE void load_scalars(const LgarDeviceState& d, int64_t col, LgScalars& s) {
  const double* p = d.scal + col;
  const int64_t st = d.stride;
  s.volon = p[LGAR_S_VOLON * st];
  s.precip_prev = p[LGAR_S_PRECIP_PREV * st];
  s.cr_fast = p[LGAR_S_CR_FAST * st];
  s.cr_slow = p[LGAR_S_CR_SLOW * st];
  s.prev_aet = p[LGAR_S_PREV_AET * st];
  s.prev_pet = p[LGAR_S_PREV_PET * st];
  s.prev_rech = p[LGAR_S_PREV_RECH * st];
  s.acc_pet = p[LGAR_S_ACC_PET * st];
  s.acc_fd = p[LGAR_S_ACC_FD * st];
  s.timestep_h = p[LGAR_S_TIMESTEP_H * st];
  s.storage = p[LGAR_S_STORAGE * st];
  s.fd_dzdt = p[LGAR_S_FD_DZDT * st];
  const int b = d.bits[col];
  s.cache_fluxes = b & 1;
  s.runoff_prev = (b >> 1) & 1;
  s.cache_count = (b >> 8) & 0xFF;
  s.status = d.status[col];
  s.rp = &d.classes[d.class_id[col]].par;
}

LG_FN_NOINLINE void store_scalars(const LgarDeviceState& d, int64_t col, const LgScalars& s, bool active) {
  double* p = d.scal + col;
  const int64_t st = d.stride;
  p[LGAR_S_VOLON * st] = s.volon;
  p[LGAR_S_PRECIP_PREV * st] = s.precip_prev;
  p[LGAR_S_CR_FAST * st] = s.cr_fast;
  p[LGAR_S_CR_SLOW * st] = s.cr_slow;
  p[LGAR_S_PREV_AET * st] = s.prev_aet;
  p[LGAR_S_PREV_PET * st] = s.prev_pet;
  p[LGAR_S_ACC_PET * st] = s.acc_pet;
  p[LGAR_S_ACC_FD * st] = s.acc_fd;
  p[LGAR_S_TIMESTEP_H * st] = s.timestep_h;
  if (active) {  // a step that replays cached fluxes leaves these three untouched
    p[LGAR_S_PREV_RECH * st] = s.prev_rech;
    p[LGAR_S_STORAGE * st] = s.storage;
    p[LGAR_S_FD_DZDT * st] = s.fd_dzdt;
  }
  d.bits[col] = (s.cache_fluxes & 1) | ((s.runoff_prev & 1) << 1) | ((s.cache_count & 0xFF) << 8);
  d.status[col] = s.status;
}

LG_FN_NOINLINE void load_fronts(const LgarDeviceState& d, int64_t col, LgFronts& f) {
  const int n = d.nfront[col];
  unsigned long long fl = d.fflags[col];
  f.n = n;
#pragma unroll 1
  for (int i = 0; i < n; i++) {
    if (i && !(i & 15)) fl = d.fflags[(int64_t)(i >> 4) * d.stride + col];
    const int64_t o = (int64_t)i * d.stride + col;
    f.depth[i] = d.depth[o];
    f.theta[i] = d.theta[o];
    f.psi[i] = d.psi[o];
    f.K[i] = d.K[o];
    f.dzdt[i] = d.dzdt[o];
    const int nib = (int)((fl >> (4 * (i & 15))) & 0xF);
    f.layer[i] = (signed char)(nib & 7);
    f.tb[i] = (signed char)(nib >> 3);
  }
}

LG_FN_NOINLINE void store_fronts(const LgarDeviceState& d, int64_t col, const LgFronts& f) {
  unsigned long long fl = 0ull;
#pragma unroll 1
  for (int i = 0; i < f.n; i++) {
    const int64_t o = (int64_t)i * d.stride + col;
    d.depth[o] = f.depth[i];
    d.theta[o] = f.theta[i];
    d.psi[o] = f.psi[i];
    d.K[o] = f.K[i];
    d.dzdt[o] = f.dzdt[i];
    if (i && !(i & 15)) {
      d.fflags[(int64_t)((i >> 4) - 1) * d.stride + col] = fl;
      fl = 0ull;
    }
    fl |= (unsigned long long)((f.layer[i] & 7) | ((f.tb[i] ? 1 : 0) << 3)) << (4 * (i & 15));
  }
  d.fflags[(int64_t)(f.n > 0 ? (f.n - 1) >> 4 : 0) * d.stride + col] = fl;
  d.nfront[col] = f.n;
}

// GIUH convolution (giuh/giuh.c:12-39; input in cm, quirk Q12), outputs in metres
// (bmi_lgar.cxx:808-893) and the cumulative volumes of the global balance (lgar.cxx:1162-1185).
LG_FN_NOINLINE void epilogue(const LgarConfig& cfg, const LgarDeviceState& d, const LgarSinks& sinks, int64_t col,
                    const LgStepVols& v, int status, bool write_out = true) {
  const int64_t st = d.stride;
  if (status & LGAR_FAIL_MASK) {
    const double nan = LG_NAN;
#pragma unroll
    for (int k = 0; k < LGAR_NOUT; k++) {
      if (write_out) d.out[k * st + col] = nan;
      if (sinks.p[k]) sinks.p[k][col] = nan;
    }
    return;
  }
  const int N = cfg.num_giuh;
  const double r = v.volrunoff + v.volQ_CR;
  // queue slot N-1 is always zero between steps, so only N-1 slots are stored
  double carry = (N > 1) ? d.giuhq[col] : 0.0;
  const double now = carry + cfg.giuh[0] * r;
  for (int i = 1; i < N; i++) {
    double qi = (i < N - 1) ? d.giuhq[(int64_t)i * st + col] : 0.0;
    qi += cfg.giuh[i] * r;
    d.giuhq[(int64_t)(i - 1) * st + col] = qi;  // shifted left; i-1 <= N-2
  }
  const double cm_to_m = 0.01;
  double o[LGAR_NOUT];
  o[LGAR_OUT_PRECIP] = v.precip * cm_to_m;
  o[LGAR_OUT_PET] = v.PET * cm_to_m;
  o[LGAR_OUT_AET] = v.AET * cm_to_m;
  o[LGAR_OUT_RUNOFF] = v.volrunoff * cm_to_m;
  o[LGAR_OUT_GIUH_RUNOFF] = now * cm_to_m;
  o[LGAR_OUT_STORAGE] = v.volend * cm_to_m;
  o[LGAR_OUT_Q] = now * cm_to_m;
  o[LGAR_OUT_INFIL] = v.volin * cm_to_m;
  o[LGAR_OUT_PERC] = v.volrech * cm_to_m;
  o[LGAR_OUT_Q_CR] = v.volQ_CR * cm_to_m;
  o[LGAR_OUT_MBAL] = v.local_mb * cm_to_m;
  o[LGAR_OUT_CR_STORAGE] = v.volCRend * cm_to_m;
#pragma unroll
  for (int k = 0; k < LGAR_NOUT; k++) {
    if (write_out) d.out[k * st + col] = o[k];
    if (sinks.p[k]) sinks.p[k][col] = o[k];
  }
  double* cu = d.cum + col;
  cu[LGAR_CUM_PRECIP * st] += v.precip;
  cu[LGAR_CUM_IN * st] += v.volin;
  cu[LGAR_CUM_AET * st] += v.AET;
  cu[LGAR_CUM_RECH * st] += v.volrech;
  cu[LGAR_CUM_RUNOFF * st] += v.volrunoff;
  cu[LGAR_CUM_Q * st] += now;
  cu[LGAR_CUM_Q_CR * st] += v.volQ_CR;
  cu[LGAR_CUM_PET * st] += v.PET;
  cu[LGAR_CUM_GIUH * st] += now;
  cu[LGAR_CUM_RUNOFF_CR * st] += v.volQ_CR;
}

LG_FN void zero_vols(LgStepVols& v) {
  v.precip = v.PET = v.AET = v.volin = v.volrech = v.volrunoff = v.volQ_CR = 0.0;
  v.volend = v.volCRend = v.local_mb = 0.0;
}

// Clamp / PET_affects_precip, bmi_lgar.cxx:325-348.  NB the sub-step and cache decisions above it
// in the reference see the raw forcing; only the sub-cycle loop sees the adjusted one.
LG_FN void adjust_forcing(const LgarConfig& cfg, double& P, double& E, int& status) {
  if (P < 0.0) {
    P = 0.0;
    status |= LGAR_WARN_NEG_FORCING;
  }
  if (E < 0.0) {
    E = 0.0;
    status |= LGAR_WARN_NEG_FORCING;
  }
  if (cfg.PET_affects_precip) {
    if (P > E) {
      P = P - E;
      E = 0.0;
    } else {
      E = E - P;
      P = 0.0;
    }
  }
}

// Invalid soil type: Q = precip (bmi_lgar.cxx:145-179)
LG_FN void step_invalid(const LgarDeviceState& d, const LgarSinks& sinks, int64_t col, double P, bool write_out = true) {
  const int64_t st = d.stride;
  const double pm = P * 0.001;
#pragma unroll
  for (int k = 0; k < LGAR_NOUT; k++) {
    double val = (k == LGAR_OUT_PRECIP || k == LGAR_OUT_RUNOFF || k == LGAR_OUT_Q) ? pm : 0.0;
    if (write_out) d.out[k * st + col] = val;
    if (sinks.p[k]) sinks.p[k][col] = val;
  }
  double* cu = d.cum + col;
  cu[LGAR_CUM_PRECIP * st] += P * 0.1;
  cu[LGAR_CUM_RUNOFF * st] += P * 0.1;
  cu[LGAR_CUM_Q * st] += P * 0.1;
}

// Body of lgar_step_stream_kernel for one column.  Returns true when the column must compute fluxes
// this step (and was therefore left untouched for the active kernel).
LG_FN bool lg_stream_column(const LgarConfig& cfg, const LgarDeviceState& d, const LgarSinks& sinks, int64_t col, int force_all_active) {
  const LgarColumnClass* cls = d.classes + d.class_id[col];
  double P = d.precip[col], E = d.pet[col];
  if (cls->invalid_soil_type) {
    step_invalid(d, sinks, col, P);
    return false;
  }
  LgScalars s;
  load_scalars(d, col, s);
  if (s.status & LGAR_FAIL_MASK) {
    LgStepVols v;
    zero_vols(v);
    epilogue(cfg, d, sinks, col, v, s.status);
    return false;
  }
  const LgPrologue pro = lg_prologue(cfg, s, P);
  if (!pro.cache_fluxes || force_all_active) return true;
  if (cfg.adaptive_timestep) s.timestep_h = pro.subtimestep_h;
  s.cache_fluxes = 1;
  s.cache_count = pro.cache_count;
  adjust_forcing(cfg, P, E, s.status);
  LgStepVols v;
  zero_vols(v);
  if (d.paths) { LgCtx pc; pc.paths = d.paths; LG_PATH(pc, LGAR_PATH_CACHED_STEP); }
  lg_step_cached(cfg, s, pro, P, E, v);
  if (!(fabs(v.local_mb) <= cfg.mbal_tol)) s.status |= (isnan(v.local_mb) ? LGAR_FAIL_NAN : LGAR_FAIL_MBAL);
  store_scalars(d, col, s, false);
  epilogue(cfg, d, sinks, col, v, s.status);
  return false;
}

// Body of lgar_step_active_kernel for one work-list entry: the full Update() of one column.
LG_FN_NOINLINE void lg_active_column(const LgarConfig& cfg, const LgarDeviceState& d, const LgarSinks& sinks, int64_t col, bool write_out = true) {
  const LgarColumnClass* cls = d.classes + d.class_id[col];
  double P = d.precip[col], E = d.pet[col];
  LgScalars s;
  load_scalars(d, col, s);
  LgPrologue pro = lg_prologue(cfg, s, P);
  if (cfg.adaptive_timestep) s.timestep_h = pro.subtimestep_h;
  s.cache_count = pro.cache_count;
  adjust_forcing(cfg, P, E, s.status);
  LgStepVols v;
  zero_vols(v);
  if (pro.cache_fluxes) {  // only reachable with force_all_active (the two kernels tested against each other)
    s.cache_fluxes = 1;
    lg_step_cached(cfg, s, pro, P, E, v);
    if (!(fabs(v.local_mb) <= cfg.mbal_tol)) s.status |= (isnan(v.local_mb) ? LGAR_FAIL_NAN : LGAR_FAIL_MBAL);
    store_scalars(d, col, s, false);
  } else {
    LgFronts f;
    load_fronts(d, col, f);
    s.cache_fluxes = 0;
    // SFT coupling: this column's frozen factors of the step (set by the engine from soil_temperature_profile)
    double ff[LGAR_MAXL + 2];
    const double* ffp = lg_ff_ones;
    if (d.ff) {
      for (int k = 0; k <= LGAR_MAXL; k++) ff[k] = d.ff[(int64_t)k * d.stride + col];
      ffp = ff;
    }
    lg_step_active(cfg, *cls, s, f, pro, P, E, v, ffp, d.paths, d.use_ff != 0);
    store_scalars(d, col, s, true);
    store_fronts(d, col, f);
  }
  epilogue(cfg, d, sinks, col, v, s.status, write_out);
}

// ------------------------------------------------------------------------------------------------
// Window stepping: K forcing hours per launch with per-column time pointers (lgar_window_kernel).
// ------------------------------------------------------------------------------------------------
LG_FN LgarSinks lg_sinks_at(const LgarWindow& w, int t) {
  LgarSinks s;
#pragma unroll
  for (int k = 0; k < LGAR_NOUT; k++) s.p[k] = w.sinks.p[k] ? w.sinks.p[k] + (int64_t)t * w.sink_stride : nullptr;
  return s;
}

// GIUH queue and cumulative volumes of one column held in registers while the column is walked through
// many hours (same arithmetic, same order as epilogue()).  Up to LGAR_EPI_REGS ordinates live in
// statically indexed registers; longer unit hydrographs use the thread-private array qm.  (The first
// version indexed one array dynamically: 1 KB of local-memory traffic per column-hour, which made the
// FETCH kernel L2-write bound -- profiles/r1_ncu_summary.md r1f.)
#define LGAR_EPI_REGS 8
struct alignas(16) LgEpi {
  double qr[LGAR_EPI_REGS];  // GIUH queue slots 0..N-2 when N <= LGAR_EPI_REGS (longer queues stay in d.giuhq)
  double cum[LGAR_NCUM];
};
LG_FN void lg_epi_load(const LgarConfig& cfg, const LgarDeviceState& d, int64_t col, LgEpi& e) {
  const int64_t st = d.stride;
  const int N = cfg.num_giuh;
#pragma unroll
  for (int i = 0; i < LGAR_EPI_REGS; i++) e.qr[i] = (N <= LGAR_EPI_REGS && i + 1 < N) ? d.giuhq[(int64_t)i * st + col] : 0.0;
#pragma unroll
  for (int k = 0; k < LGAR_NCUM; k++) e.cum[k] = d.cum[(int64_t)k * st + col];
}
LG_FN void lg_epi_store(const LgarConfig& cfg, const LgarDeviceState& d, int64_t col, const LgEpi& e) {
  const int64_t st = d.stride;
  const int N = cfg.num_giuh;
  if (N <= LGAR_EPI_REGS) {
#pragma unroll
    for (int i = 0; i < LGAR_EPI_REGS - 1; i++)
      if (i + 1 < N) d.giuhq[(int64_t)i * st + col] = e.qr[i];
  }
#pragma unroll
  for (int k = 0; k < LGAR_NCUM; k++) d.cum[(int64_t)k * st + col] = e.cum[k];
}
// one step of the epilogue at hour t of window w (sinks addressed at [t][col])
LG_FN void lg_epi_step(const LgarConfig& cfg, const LgarDeviceState& d, const LgarWindow& w, int t, int64_t col, LgEpi& e, const LgStepVols& v,
                       int status, bool write_out) {
  const int64_t st = d.stride;
  const int64_t so = (int64_t)t * w.sink_stride + col;
  if (status & LGAR_FAIL_MASK) {
    const double nan = LG_NAN;
    if (write_out || w.has_sinks) {
#pragma unroll
      for (int k = 0; k < LGAR_NOUT; k++) {
        if (write_out) d.out[k * st + col] = nan;
        if (w.sinks.p[k]) w.sinks.p[k][so] = nan;
      }
    }
    return;
  }
  const int N = cfg.num_giuh;
  const double r = v.volrunoff + v.volQ_CR;
  double now;
  if (N <= LGAR_EPI_REGS) {
    const double carry = (N > 1) ? e.qr[0] : 0.0;
    now = carry + cfg.giuh[0] * r;
#pragma unroll
    for (int i = 1; i < LGAR_EPI_REGS; i++) {
      if (i < N) {
        double qi = (i < N - 1) ? e.qr[i] : 0.0;
        qi += cfg.giuh[i] * r;
        e.qr[i - 1] = qi;
      }
    }
  } else {  // long unit hydrograph: the queue is read-modify-written in HBM, as in epilogue()
    const double carry = d.giuhq[col];
    now = carry + cfg.giuh[0] * r;
    for (int i = 1; i < N; i++) {
      double qi = (i < N - 1) ? d.giuhq[(int64_t)i * st + col] : 0.0;
      qi += cfg.giuh[i] * r;
      d.giuhq[(int64_t)(i - 1) * st + col] = qi;
    }
  }
  e.cum[LGAR_CUM_PRECIP] += v.precip;
  e.cum[LGAR_CUM_IN] += v.volin;
  e.cum[LGAR_CUM_AET] += v.AET;
  e.cum[LGAR_CUM_RECH] += v.volrech;
  e.cum[LGAR_CUM_RUNOFF] += v.volrunoff;
  e.cum[LGAR_CUM_Q] += now;
  e.cum[LGAR_CUM_Q_CR] += v.volQ_CR;
  e.cum[LGAR_CUM_PET] += v.PET;
  e.cum[LGAR_CUM_GIUH] += now;
  e.cum[LGAR_CUM_RUNOFF_CR] += v.volQ_CR;
  if (!(write_out || w.has_sinks)) return;
  const double cm_to_m = 0.01;
  double o[LGAR_NOUT];
  o[LGAR_OUT_PRECIP] = v.precip * cm_to_m;
  o[LGAR_OUT_PET] = v.PET * cm_to_m;
  o[LGAR_OUT_AET] = v.AET * cm_to_m;
  o[LGAR_OUT_RUNOFF] = v.volrunoff * cm_to_m;
  o[LGAR_OUT_GIUH_RUNOFF] = now * cm_to_m;
  o[LGAR_OUT_STORAGE] = v.volend * cm_to_m;
  o[LGAR_OUT_Q] = now * cm_to_m;
  o[LGAR_OUT_INFIL] = v.volin * cm_to_m;
  o[LGAR_OUT_PERC] = v.volrech * cm_to_m;
  o[LGAR_OUT_Q_CR] = v.volQ_CR * cm_to_m;
  o[LGAR_OUT_MBAL] = v.local_mb * cm_to_m;
  o[LGAR_OUT_CR_STORAGE] = v.volCRend * cm_to_m;
#pragma unroll
  for (int k = 0; k < LGAR_NOUT; k++) {
    if (write_out) d.out[k * st + col] = o[k];
    if (w.sinks.p[k]) w.sinks.p[k][so] = o[k];
  }
}

// Advance one column from hour t through every step that needs no flux computation (cached-flux
// replay, invalid soil, failed column), scalars held in registers across hours.  Returns the hour at
// which the column needs an active step, or w.K when it reached the end of the window.
LG_FN_NOINLINE int lg_advance_column(const LgarConfig& cfg, const LgarDeviceState& d, const LgarWindow& w, int64_t col, int t,
                            unsigned long long* front_sum) {
  const LgarColumnClass* cls = d.classes + d.class_id[col];
  const int K = w.K;
  if (cls->invalid_soil_type) {
    for (; t < K; t++) {
      const LgarSinks sk = lg_sinks_at(w, t);
      step_invalid(d, sk, col, w.precip[(int64_t)t * w.series_stride + col], t == K - 1);
      if (w.nf_sink) w.nf_sink[(int64_t)t * w.sink_stride + col] = 0;
    }
    return K;
  }
  LgScalars s;
  load_scalars(d, col, s);
  const int nf = (w.nf_sink || front_sum) ? d.nfront[col] : 0;
  bool dirty = false, epi_loaded = false;
  LgEpi ep;
  for (; t < K; t++) {
    double P = w.precip[(int64_t)t * w.series_stride + col], E = w.pet[(int64_t)t * w.series_stride + col];
    LgStepVols v;
    zero_vols(v);
    if (!(s.status & LGAR_FAIL_MASK)) {
      const LgPrologue pro = lg_prologue_inl(cfg, s, P);
      if (!pro.cache_fluxes || w.force_all_active) break;
      if (cfg.adaptive_timestep) s.timestep_h = pro.subtimestep_h;
      s.cache_fluxes = 1;
      s.cache_count = pro.cache_count;
      adjust_forcing(cfg, P, E, s.status);
      lg_step_cached_inl(cfg, s, pro, P, E, v);
      if (!(fabs(v.local_mb) <= cfg.mbal_tol)) s.status |= (isnan(v.local_mb) ? LGAR_FAIL_NAN : LGAR_FAIL_MBAL);
      dirty = true;
    }
    if (!epi_loaded) {
      lg_epi_load(cfg, d, col, ep);
      epi_loaded = true;
    }
    lg_epi_step(cfg, d, w, t, col, ep, v, s.status, t == K - 1);
    if (w.nf_sink) w.nf_sink[(int64_t)t * w.sink_stride + col] = nf;
    if (front_sum) *front_sum += (unsigned)nf;
  }
  if (epi_loaded) lg_epi_store(cfg, d, col, ep);
  if (dirty) store_scalars(d, col, s, false);
  return t;
}

// One hour of that walk for a resident column: false if hour t needs the full Update() (nothing has been changed then),
// true if it was finished from the scalars by replaying the cached fluxes (bmi_lgar.cxx:260-318,683-695).
LG_FN bool lg_resident_hour(const LgarConfig& cfg, const LgarDeviceState& d, const LgarWindow& w, int64_t col, int t, int nf, LgScalars& s,
                            LgEpi& ep, unsigned long long* front_sum, double P, double E) {
  LgStepVols v;
  zero_vols(v);
  if (!(s.status & LGAR_FAIL_MASK)) {
    const LgPrologue pro = lg_prologue_inl(cfg, s, P);
    if (!pro.cache_fluxes || w.force_all_active) return false;
    if (cfg.adaptive_timestep) s.timestep_h = pro.subtimestep_h;
    s.cache_fluxes = 1;
    s.cache_count = pro.cache_count;
    adjust_forcing(cfg, P, E, s.status);
    if (w.paths) { LgCtx pc; pc.paths = w.paths; LG_PATH(pc, LGAR_PATH_CACHED_STEP); }
    lg_step_cached_inl(cfg, s, pro, P, E, v);
    if (!(fabs(v.local_mb) <= cfg.mbal_tol)) s.status |= (isnan(v.local_mb) ? LGAR_FAIL_NAN : LGAR_FAIL_MBAL);
  }
  lg_epi_step(cfg, d, w, t, col, ep, v, s.status, t == w.K - 1);
  if (w.nf_sink) w.nf_sink[(int64_t)t * w.sink_stride + col] = nf;
  if (front_sum) *front_sum += (unsigned)nf;
  return true;
}

// The same walk for a column whose scalars, GIUH queue and cumulative volumes are resident in its lane
// record (lgar_lanes.cuh): nothing is read from or written to the SoA state.
// forcing hours of the window that are present (LgarWindow::frontier1)
LG_FN int lg_hours_present(const LgarWindow& w) { return w.frontier1 > 0 ? w.frontier1 - 1 : w.K; }
LG_FN int lg_advance_resident(const LgarConfig& cfg, const LgarDeviceState& d, const LgarWindow& w, const LgarColumnClass* cls, int64_t col, int t,
                              int nf, LgScalars& s, LgEpi& ep, unsigned long long* front_sum) {
  const int K = lg_hours_present(w);  // the walk ends where the forcing ends
  if (cls->invalid_soil_type) {  // Q = precip (bmi_lgar.cxx:145-179); cumulative volumes as step_invalid()
    for (; t < K; t++) {
      const double P = w.precip[(int64_t)t * w.series_stride + col];
      const double pm = P * 0.001;
      const int64_t so = (int64_t)t * w.sink_stride + col;
#pragma unroll
      for (int k = 0; k < LGAR_NOUT; k++) {
        const double val = (k == LGAR_OUT_PRECIP || k == LGAR_OUT_RUNOFF || k == LGAR_OUT_Q) ? pm : 0.0;
        if (t == w.K - 1) d.out[k * d.stride + col] = val;
        if (w.sinks.p[k]) w.sinks.p[k][so] = val;
      }
      ep.cum[LGAR_CUM_PRECIP] += P * 0.1;
      ep.cum[LGAR_CUM_RUNOFF] += P * 0.1;
      ep.cum[LGAR_CUM_Q] += P * 0.1;
      if (w.nf_sink) w.nf_sink[so] = 0;
    }
    return K;
  }
  for (; t < K; t++) {
    const double P = w.precip[(int64_t)t * w.series_stride + col], E = w.pet[(int64_t)t * w.series_stride + col];
    if (!lg_resident_hour(cfg, d, w, col, t, nf, s, ep, front_sum, P, E)) break;
  }
  return t;
}

// The full Update() of one column at hour t of the window.
LG_FN void lg_active_column_at(const LgarConfig& cfg, const LgarDeviceState& d, const LgarWindow& w, int64_t col, int t,
                               unsigned long long* front_sum) {
  LgarDeviceState dd = d;
  dd.precip = w.precip + (int64_t)t * w.series_stride;
  dd.pet = w.pet + (int64_t)t * w.series_stride;
  const LgarSinks sk = lg_sinks_at(w, t);
  lg_active_column(cfg, dd, sk, col, t == w.K - 1);
  if (w.nf_sink || front_sum) {
    const int nf = d.nfront[col];
    if (w.nf_sink) w.nf_sink[(int64_t)t * w.sink_stride + col] = nf;
    if (front_sum) *front_sum += (unsigned)nf;
  }
}

#endif  // LGAR_STEP_CUH
