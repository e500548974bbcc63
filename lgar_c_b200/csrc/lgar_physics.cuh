// lgar_c_b200/csrc/lgar_physics.cuh
//
// Device-side LGAR physics for ONE soil column held in thread-private arrays: the body of the
// batched step kernels in lgar_kernels.cu.  It reproduces, step for step, what the reference does
// under BmiLGAR::Update() (reference src/bmi_lgar.cxx:133-895 and the functions of src/lgar.cxx,
// src/soil_funcs.cxx, src/aet.cxx, src/conceptual_reservoir.cxx, giuh/giuh.c it calls) because
// the ad-hoc iterative searches only define their answer through their iteration sequence
// (SURVEY.md section 7 "step-faithfulness").  What is different is everything around the
// arithmetic: no heap, no list nodes, no strings -- fronts are slots of fixed arrays, the front
// number is the slot index, soil constants come pre-derived from a read-only class row, and a
// failure sets a status bit instead of aborting the process.
//
// The header is compiled for the device by nvcc (LG_FN = __device__).  tests/hostsim compiles the
// same text with a host compiler (LGAR_HOSTSIM) purely to debug kernel logic against the oracle on
// a machine without a GPU; that build is test infrastructure and is never part of liblgarb.so.
#ifndef LGAR_PHYSICS_CUH
#define LGAR_PHYSICS_CUH

#include "lgar_types.h"
#include "lgar_pow.cuh"

#ifdef LGAR_HOSTSIM
#include <cmath>
#define LG_FN static inline
#define LG_FN_NOINLINE static
#define LG_LANES_HERE() 0u
#define LG_RECONVERGE(m) ((void)(m))
using std::exp; using std::fabs; using std::fmax; using std::fmin; using std::isinf; using std::isnan; using std::log; using std::sqrt;
#else
#define LG_FN __device__ __forceinline__
#ifdef LGAR_TU_LOCAL   // a second translation unit that includes the physics (lgar_finish_smem.cu): internal linkage
#define LG_FN_NOINLINE static __device__ __noinline__
#else
#define LG_FN_NOINLINE __device__ __noinline__
#endif
// Lanes that entered a function together meet again at LG_RECONVERGE before the expensive common part.  Without it
// a function with an early exit in a divergent loop runs its common tail once per group of lanes (ncu source
// view of the TAIL kernel: the K/psi refresh of lg_move_end ran at 17 of 32 lanes before its own loop divergence).
// Every lane of LG_LANES_HERE() must reach the LG_RECONVERGE: no return in between.
#define LG_LANES_HERE() __activemask()
#define LG_RECONVERGE(m) __syncwarp(m)
#endif

// reference src/lgar.cxx:63-70
#define LG_THRESHOLD_NO_MOISTURE_DIFF 1.E-15
#define LG_MBAL_TOL 1.E-10
#define LG_MAX_ITER_MBAL 100000
#define LG_MAX_ITER_SAT 10000
#define LG_TRUNCATION_DEPTH 1.E-9
#define LG_FACTOR_LIMITS_CROSSING 2.0
#define LG_DEPTH_AVOIDS_SAME 1.E-6
#define LG_PSI_UPPER_LIM 1.E7
// reference src/bmi_lgar.cxx:22-57
#define LG_PRECIP_THRESHOLD_MM_PER_H 1.0e-6
#define LG_PET_EPSILON 1.0e-8
#define LG_AET_PET_RATIO_THRESHOLD 0.75
#define LG_VOLON_THRESHOLD_CM 1.0e-16
#define LG_BOTTOM_FLUX_THRESHOLD_CM 1.0e-4
#define LG_THRESHOLD_DZDT 1.0e0
#define LG_CACHE_RESET_STEPS 24
#define LG_SMALL_EPS 1.E-12

struct LgFronts {
  double depth[LGAR_W], theta[LGAR_W], psi[LGAR_W], K[LGAR_W], dzdt[LGAR_W];
  signed char layer[LGAR_W];
  signed char tb[LGAR_W];
  int n;
};

// state_previous (bmi_lgar.cxx:370-374): only depth/theta/psi are ever read back from it
struct LgOld {
  double depth[LGAR_W], theta[LGAR_W], psi[LGAR_W];
  int n;
};

struct LgScalars {
  double volon, precip_prev, cr_fast, cr_slow, prev_aet, prev_pet, prev_rech, acc_pet, acc_fd;
  double timestep_h, storage, fd_dzdt;
  int cache_fluxes, cache_count, runoff_prev;
  int status;
  const LgarColumnParams* rp;  // the column's reservoir / preferential-flow / ponding parameters (class row), set by load_scalars
};

struct LgCtx {
  const LgarConfig* cfg;
  const LgarColumnClass* cls;
  LgFronts* f;
  int status;
  // frozen_factor[layer] of the SFT coupling (1-based; reference lgar.cxx:1109-1149), all 1.0 unless the run is
  // coupled.  The places that use it -- and the three that do not (quirk Q11) -- follow the reference one by one;
  // x * 1.0 == x, so an uncoupled run is bit for bit what it was without the factor.
  const double* ff;
  unsigned long long* paths;  // null, or [LGAR_NPATH] counters of the rare paths taken (lgarb_set_path_counters)
  bool use_ff;                // psi searches take the fast path of lgar_search.cuh (a latency tool, see there)
};
// The searches that run INSIDE a step function (hour-by-hour kernels; the rare correction search of a layer crossing at the domain
// bottom inside TAIL) walk the plain way on the device: the fast path's registers would be charged to every kernel that can reach
// it (SEARCH 115 -> 168, CORR 62 -> 146 registers when it was reachable from them unconditionally).  The host build of the kernel
// text honours the flag, so the CPU suite exercises both.
#ifdef LGAR_HOSTSIM
#define LG_STRAIGHT_FF(ctx) ((ctx).use_ff)
#else
#define LG_STRAIGHT_FF(ctx) false
#endif
#ifdef LGAR_HOSTSIM
#define LG_PATH(ctx, id) do { if ((ctx).paths) (ctx).paths[id]++; } while (0)
#else
#define LG_PATH(ctx, id) do { if ((ctx).paths) atomicAdd((ctx).paths + (id), 1ULL); } while (0)
#endif
#ifdef LGAR_HOSTSIM
static const double lg_ff_ones[LGAR_MAXL + 2] = {1.0, 1.0, 1.0, 1.0, 1.0, 1.0};
#elif defined(LGAR_TU_LOCAL)
static __device__ const double lg_ff_ones[LGAR_MAXL + 2] = {1.0, 1.0, 1.0, 1.0, 1.0, 1.0};
#else
__device__ const double lg_ff_ones[LGAR_MAXL + 2] = {1.0, 1.0, 1.0, 1.0, 1.0, 1.0};
#endif

// ------------------------------------------------------------------------------------------------
// L1: van Genuchten constitutive relations, reference src/soil_funcs.cxx
// ------------------------------------------------------------------------------------------------
LG_FN double lg_sq(double x) { return x * x; }

// x^y.  The reference calls libm pow() everywhere (soil_funcs.cxx:147-187, aet.cxx:60, conceptual_reservoir.cxx:19-35), and the
// engine returns the same double: lgar_pow.cuh restates the host library's table-driven algorithm operation for operation
// (bit-identical to glibc's pow on 3e8 samples, tests/test_pow_exact.py), so that what separated the engine from the reference
// in round 1 -- exp(y log x), a few ulp from pow in most calls, enough to end a psi search one trip apart --
// is gone.
//
// Build variants (measurement and cross-checks only):
//   -DLGAR_FAST_POW      round 1's exp(y log x) with log(psi) shared by the layers of a search trip
//   -DLGAR_ACCURATE_POW  the library pow of the compiling toolchain (CUDA's on the device; glibc's in the host build of the
//                        kernel text, tests/hostsim -- the function the reference itself calls)
//
// Code size matters as much as instruction count here: with every soil function inlined the lanes kernel of round 1 was 437 KB
// of SASS and warps, each at a different place of the column program, stalled 26 of 34 cycles per instruction on instruction
// fetch (profiles/r1_ncu_summary.md).  So pow exists ONCE as an out-of-line function and the soil functions are out of line
// too; only the psi-search iteration inlines the common case (lgp_try_pow) and calls the full function for the rest.
#if defined(LGAR_FAST_POW)
LG_FN_NOINLINE double lg_exp(double x) { return exp(x); }
LG_FN_NOINLINE double lg_log(double x) { return log(x); }
LG_FN double lg_pow(double x, double y) { return lg_exp(y * lg_log(x)); }
LG_FN double lg_pow_hot(double x, double y) { return lg_pow(x, y); }
// calc_theta_from_h with log(h) supplied, fully inlined: the psi search evaluates every layer at the same psi
LG_FN double lg_theta_from_lnh(double lnh, const LgarLayer& s) {
  const double y = exp(s.n * (s.ln_alpha + lnh));       // (alpha*h)^n
  return exp(-s.m * log(1.0 + y)) * s.de + s.theta_r;   // 1/(1+y)^m * (theta_e-theta_r) + theta_r
}
LG_FN double lg_lnh(double h) { return log(h); }
// calc_theta_from_h, soil_funcs.cxx:147-150
LG_FN_NOINLINE double lg_theta_from_h(double h, const LgarLayer& s) {
  const double y = lg_exp(s.n * (s.ln_alpha + lg_log(h)));
  return lg_exp(-s.m * lg_log(1.0 + y)) * s.de + s.theta_r;
}
// calc_Se_from_h, soil_funcs.cxx:155-159
LG_FN_NOINLINE double lg_Se_from_h(double h, const LgarLayer& s) {
  if (fabs(h) < 1.0E-10) return 1.0;
  return lg_exp(-s.m * lg_log(1.0 + lg_exp(s.n * (s.ln_alpha + lg_log(h)))));
}
LG_FN double lg_pow2(double x) { return x * x; }
LG_FN double lg_pow3(double x) { return x * x * x; }
#else
#if defined(LGAR_ACCURATE_POW)
#if defined(LGAR_HOSTSIM) && defined(LGAR_POWL_POW)
// host experiment (tools/bitexact_sample.py --powl): a pow of different provenance but of glibc's class of accuracy -- the x87
// long-double powl (64-bit significand) rounded to double, i.e. the correctly rounded result in all but ~1e-3 of the calls
LG_FN_NOINLINE double lg_pow(double x, double y) { return (double)powl((long double)x, (long double)y); }
#else
LG_FN_NOINLINE double lg_pow(double x, double y) { return pow(x, y); }
#endif
LG_FN double lg_pow_hot(double x, double y) { return lg_pow(x, y); }
#else
LG_FN_NOINLINE double lg_pow(double x, double y) { return lgp_pow(x, y); }
LG_FN double lg_pow_hot(double x, double y) {
  double r;
  if (lgp_try_pow(x, y, &r)) return r;
  return lg_pow(x, y);
}
#endif
// calc_theta_from_h inside the psi search ("lnh" carries h itself in these builds: every layer needs its own log(alpha h))
LG_FN double lg_theta_from_lnh(double lnh, const LgarLayer& s) {
  return 1.0 / (lg_pow_hot(1.0 + lg_pow_hot(s.alpha * lnh, s.n), s.m)) * s.de + s.theta_r;
}
LG_FN double lg_lnh(double h) { return h; }
// calc_theta_from_h, soil_funcs.cxx:147-150
LG_FN_NOINLINE double lg_theta_from_h(double h, const LgarLayer& s) {
  return 1.0 / (lg_pow(1.0 + lg_pow(s.alpha * h, s.n), s.m)) * s.de + s.theta_r;
}
// calc_Se_from_h, soil_funcs.cxx:155-159
LG_FN_NOINLINE double lg_Se_from_h(double h, const LgarLayer& s) {
  if (fabs(h) < 1.0E-10) return 1.0;
  return 1.0 / (lg_pow(1.0 + lg_pow(s.alpha * h, s.n), s.m));
}
// pow(x, 2.0) (soil_funcs.cxx:166): the reference's compiler folds a pow with the constant exponent 2.0 into x*x (GCC does so at
// every optimisation level that inlines builtins; the disassembly of calc_K_from_Se in oracle/_ref shows two calls of pow and one
// mulsd x,x), and the true pow(x, 2.0) differs from the product in the last bit now and then.  pow(x, 3.0) (aet.cxx:60) stays
// a call of pow in the reference's object code, so it is one here.
LG_FN double lg_pow2(double x) { return x * x; }
LG_FN double lg_pow3(double x) { return lg_pow(x, 3.0); }
#endif
// calc_K_from_Se, soil_funcs.cxx:164-167
// Ksat is an argument: the reference passes Ks * frozen_factor[layer] at some call sites and plain Ks at others
LG_FN_NOINLINE double lg_K_from_Se(double Se, const LgarLayer& s, double Ksat) {
  return Ksat * sqrt(Se) * lg_pow2(1.0 - lg_pow(1.0 - lg_pow(Se, s.inv_m), s.m));
}
// calc_h_from_Se, soil_funcs.cxx:172-179
LG_FN_NOINLINE double lg_h_from_Se(double Se, const LgarLayer& s) {
  double r = s.inv_alpha * lg_pow(lg_pow(Se, s.neg_inv_m) - 1.0, s.inv_n);
  if (r > 1.E20) r = 1.E20;
  return r;
}
// calc_Se_from_theta, soil_funcs.cxx:184-187
LG_FN double lg_Se_from_theta(double theta, const LgarLayer& s) { return (theta - s.theta_r) / s.de; }

// calc_Geff, soil_funcs.cxx:34-140.  Closed form: Se_f from theta1, Se_i from theta2.
LG_FN_NOINLINE double lg_Geff(const LgarConfig& cfg, double theta1, double theta2, const LgarLayer& s, double Ksat) {
  if (!cfg.use_closed_form_G) {
    double Se_i = lg_Se_from_theta(theta1, s);
    double Se_f = lg_Se_from_theta(theta2, s);
    double h_i = lg_h_from_Se(Se_i, s);
    double h_f = lg_h_from_Se(Se_f, s);
    double dh = (h_i - h_f) / (double)cfg.nint;
    dh = dh * 0.01;
    double Geff = 0.0;
    double K1 = lg_K_from_Se(Se_i, s, Ksat);
    double h2 = h_f + dh;
    while (h2 < h_i) {
      double Se2 = lg_Se_from_h(h2, s);
      double K2 = lg_K_from_Se(Se2, s, Ksat);
      if ((K1 - K2) / K2 > 0.02)
        dh = dh * 0.5;
      else
        dh = dh * 10.0;
      Geff += (K1 + K2) * dh / 2.0;
      K1 = K2;
      h2 += dh;
    }
    return fabs(Geff / Ksat);
  } else {
    double Se_f = lg_Se_from_theta(theta1, s);
    double Se_i = lg_Se_from_theta(theta2, s);
    double pf = lg_pow(Se_f, s.p_exp);
    double Geff = s.H_c * (lg_pow(Se_i, s.p_exp) - pf) / (1 - pf);
    if (isinf(Geff)) Geff = s.H_c;
    if (isnan(Geff)) Geff = s.H_c;
    return Geff;
  }
}

#include "lgar_search.cuh"

// which exit ended an iterated psi search (instrumentation: the conditions are those of the loop, in its order)
LG_FN void lg_path_search_exit(LgCtx& c, const LgSearch& q) {
  if (!c.paths) return;
  LG_PATH(c, LGAR_PATH_SEARCH);
  int id;
  if (!(q.delta_mass > LG_MBAL_TOL)) id = LGAR_PATH_SEARCH_CONVERGED;
  else if (lg_digit_brk_first(q)) id = LGAR_PATH_SEARCH_STEP_EXIT;
  else if (q.count_no_mass_change == 5 && q.precip_mass_to_add < 1.E-12) id = LGAR_PATH_SEARCH_NOCHANGE;
  else if (q.iter > LG_MAX_ITER_MBAL) id = LGAR_PATH_SEARCH_MAXITER;
  else if ((q.psi_loc > LG_PSI_UPPER_LIM) && (q.iter > LG_MAX_ITER_MBAL / 100)) id = LGAR_PATH_SEARCH_PSI_UPPER;
  else id = LGAR_PATH_SEARCH_PSI_ZERO;
  LG_PATH(c, id);
}

// ------------------------------------------------------------------------------------------------
// L0: list behaviour on fixed slots, reference src/linked_list.cxx
// ------------------------------------------------------------------------------------------------
LG_FN void lg_copy_slot(LgFronts& f, int dst, int src) {
  f.depth[dst] = f.depth[src];
  f.theta[dst] = f.theta[src];
  f.psi[dst] = f.psi[src];
  f.K[dst] = f.K[src];
  f.dzdt[dst] = f.dzdt[src];
  f.layer[dst] = f.layer[src];
  f.tb[dst] = f.tb[src];
}

// listDeleteFront, linked_list.cxx:208-293 (psi/theta/K re-propagation :267-288; skipped for the head)
LG_FN_NOINLINE int lg_front_delete(LgCtx& c, int idx) {
  LgFronts& f = *c.f;
  if (idx < 0 || idx >= f.n) {
    c.status |= LGAR_FAIL_LIST;
    return idx;
  }
  if (idx == f.n - 1) c.status |= LGAR_FAIL_LIST;
  for (int i = idx; i < f.n - 1; i++) lg_copy_slot(f, i, i + 1);
  f.n--;
  if (idx != 0) {
    for (int wf = f.n - 1; wf != 0; wf--) {
      int i = wf - 1;
      if (f.tb[i]) {
        const LgarLayer& sp = c.cls->lay[f.layer[i]];
        f.psi[i] = f.psi[wf];
        f.theta[i] = lg_theta_from_h(f.psi[i], sp);
        double Se = lg_Se_from_theta(f.theta[i], sp);
        f.K[i] = lg_K_from_Se(Se, sp, sp.Ksat);  // un-frozen Ks in the reference (linked_list.cxx:283, quirk Q11)
      }
    }
  }
  return idx;
}

// listInsertFirst, linked_list.cxx:97-123
LG_FN void lg_front_insert_first(LgCtx& c, double depth, double theta, int layer, int to_bottom) {
  LgFronts& f = *c.f;
  if (f.n >= LGAR_W) {
    c.status |= LGAR_FAIL_OVERFLOW;
    return;
  }
  for (int i = f.n; i > 0; i--) lg_copy_slot(f, i, i - 1);
  f.n++;
  f.depth[0] = depth;
  f.theta[0] = theta;
  f.layer[0] = (signed char)layer;
  f.tb[0] = (signed char)to_bottom;
  f.dzdt[0] = 0.0;
  f.psi[0] = 0.0;
  f.K[0] = 0.0;
}

// ------------------------------------------------------------------------------------------------
// L2: LGAR physics, reference src/lgar.cxx
// ------------------------------------------------------------------------------------------------

// lgar_calc_mass_bal, lgar.cxx:2632-2668
LG_FN_NOINLINE double lg_mass_bal(const LgCtx& c) {
  const LgFronts& f = *c.f;
  const double* cum = c.cls->cum;
  double sum = 0.0;
  // one expression for the three cases of the reference (x - 0.0 == x and a*b == b*a exactly): no branch in the loop
  // body, so the lanes of a warp stay together
  for (int i = 0; i < f.n; i++) {
    const double base = cum[f.layer[i] - 1];
    const bool same_layer_below = (i + 1 < f.n) && (f.layer[i + 1] == f.layer[i]);
    const double theta_below = same_layer_below ? f.theta[i + 1] : 0.0;
    sum += (f.depth[i] - base) * (f.theta[i] - theta_below);
  }
  return sum;
}

// wetting_front_free_drainage, lgar.cxx:1227-1254 (1-based)
LG_FN int lg_wf_free_drainage(const LgFronts& f) {
  int wf = 1;
  for (int i = 0; i + 1 < f.n; i++) {
    if (f.layer[i] == f.layer[i + 1]) break;
    wf++;
  }
  if (wf > f.n) wf--;
  return wf;
}

// calc_min_water_possible_for_free_drainage_wetting_front, lgar.cxx:3197-3232
LG_FN double lg_min_water_fd(const LgCtx& c, int wf_fd) {
  const LgFronts& f = *c.f;
  double min_storage = 0.0, previous_depth = 0.0;
  for (int i = 0; i < f.n; i++) {
    min_storage += c.cls->lay[f.layer[i]].theta_min * (f.depth[i] - previous_depth);
    if (i + 1 == wf_fd) break;
    previous_depth = f.depth[i];
  }
  return min_storage;
}
// calc_storage_in_free_drainage_wetting_front, lgar.cxx:3234-3252
LG_FN double lg_storage_fd(const LgCtx& c, int wf_fd) {
  const LgFronts& f = *c.f;
  double storage = 0.0, previous_depth = 0.0;
  for (int i = 0; i < f.n; i++) {
    storage += f.theta[i] * (f.depth[i] - previous_depth);
    if (i + 1 == wf_fd) break;
    previous_depth = f.depth[i];
  }
  return storage;
}

// The psi search of lgar_theta_mass_balance, lgar.cxx:2952-3112.  dth/dz are 1-based per layer.
LG_FN_NOINLINE double lg_theta_mass_balance(LgCtx& c, int layer_num, double psi_cm, double new_mass, double prior_mass,
                                            double precip_mass_to_add, double* AET_demand_cm, const double* dth,
                                            const double* dz) {
  const LgarColumnClass& cls = *c.cls;
  const LgarLayer& sp = cls.lay[layer_num];
  LgSearch q;
  lg_search_setup(q, layer_num, psi_cm, new_mass, prior_mass, precip_mass_to_add);
  for (int k = 1; k <= layer_num; k++) {
    q.dth[k] = dth[k];
    q.dz[k] = dz[k];
  }
  if (q.delta_mass <= LG_MBAL_TOL) return lg_theta_from_h(psi_cm, sp);
  lg_search_run(q, cls, LG_STRAIGHT_FF(c));
  lg_path_search_exit(c, q);
  if ((q.delta_mass > LG_MBAL_TOL) && (!q.wanted_to_saturate)) *AET_demand_cm = *AET_demand_cm - fabs(q.delta_mass - LG_MBAL_TOL);
  if ((q.theta >= sp.theta_e) && (q.psi_loc != 0.0)) *AET_demand_cm = 0.0;
  return q.theta;
}

// lgar_theta_mass_balance_correction, lgar.cxx:3262-3435 (front_num 1-based)
LG_FN_NOINLINE void lg_theta_mass_balance_correction(LgCtx& c, int front_num, double prior_mass) {
  LgFronts& f = *c.f;
  const LgarColumnClass& cls = *c.cls;
  int cur = front_num - 1;
  if (cur < 0 || cur >= f.n) {
    c.status |= LGAR_FAIL_LIST;
    return;
  }
  if (cur + 1 > 1) {
    if (f.tb[cur - 1]) cur = cur - 1;
  }
  while (f.tb[cur]) {
    if (cur + 1 == 1) break;
    cur = cur - 1;
    if (!f.tb[cur]) {
      cur = cur + 1;
      break;
    }
  }
  LgCorr k;
  k.cur = cur;
  k.iter = 0;
  k.new_mass = lg_mass_bal(c);
  const double psi_cm = f.psi[cur];
  k.psi_loc = psi_cm;
  k.prior_mass = prior_mass;
  k.delta_mass = fabs(k.new_mass - prior_mass);
  k.factor = fmax(1.0, psi_cm / 10.0);
  if (psi_cm > 1.E4) k.factor = psi_cm;
  k.switched = false;
  k.psi_prev = psi_cm;
  k.delta_mass_prev = k.delta_mass;
  k.count_no_mass_change = 0;
  k.aug = false;
  k.aug_extreme = false;
  k.no_ff = false;
  k.done = !(k.delta_mass > LG_MBAL_TOL);
  if (!k.done) {
    LG_PATH(c, LGAR_PATH_CORR_SEARCH);
    lg_corr_run(k, f, cls, LG_STRAIGHT_FF(c));
  }
}

// lgar_merge_wetting_fronts, lgar.cxx:1875-1953
LG_FN_NOINLINE void lg_merge(LgCtx& c) {
  LgFronts& f = *c.f;
  int cur = 0;
  for (int wf = 1; wf != f.n; wf++) {
    int nx = cur + 1, nn = cur + 2;
    if (nx >= f.n) {
      c.status |= LGAR_FAIL_LIST;
      return;
    }
    if ((f.depth[cur] > f.depth[nx]) && (f.layer[cur] == f.layer[nx]) && !f.tb[nx]) {
      if (nn >= f.n) {
        c.status |= LGAR_FAIL_LIST;
        return;
      }
      double mass_this_layer = f.depth[cur] * (f.theta[cur] - f.theta[nx]) + f.depth[nx] * (f.theta[nx] - f.theta[nn]);
      f.depth[cur] = mass_this_layer / (f.theta[cur] - f.theta[nn]);
      if (!(f.depth[cur] > 0.0)) c.status |= LGAR_FAIL_MERGE_DEPTH;
      const LgarLayer& sp = c.cls->lay[f.layer[cur]];
      double Se = lg_Se_from_theta(f.theta[cur], sp);
      f.psi[cur] = lg_h_from_Se(Se, sp);
      f.K[cur] = lg_K_from_Se(Se, sp, sp.Ksat * c.ff[f.layer[cur]]);  // lgar.cxx:1927
      lg_front_delete(c, nx);
    }
    cur = cur + 1;
    if (wf + 1 != f.n && cur >= f.n) {
      c.status |= LGAR_FAIL_LIST;
      return;
    }
  }
}

// lgar_wetting_fronts_cross_layer_boundary, lgar.cxx:1963-2167
LG_FN_NOINLINE void lg_cross_layer_boundary(LgCtx& c) {
  LgFronts& f = *c.f;
  const LgarColumnClass& cls = *c.cls;
  const double* cum = cls.cum;
  const int num_layers = cls.num_layers;
  int cur = 0;
  bool cross_necessary = false, theta_correction_necessary = false;
  int front_for_cross = -1;

  for (int wf = 1; wf != f.n; wf++) {
    int layer_num = f.layer[cur];
    int nx = cur + 1, nn = cur + 2;
    if (nx >= f.n) {
      c.status |= LGAR_FAIL_LIST;
      return;
    }
    if (f.depth[cur] > cum[layer_num] && (f.depth[nx] == cum[layer_num]) && f.tb[nx] && (layer_num != num_layers)) {
      if (nn >= f.n) {
        c.status |= LGAR_FAIL_LIST;
        return;
      }
      const LgarLayer& sp = cls.lay[layer_num];
      const LgarLayer& spn = cls.lay[layer_num + 1];
      double current_theta = fmin(sp.theta_e, f.theta[cur]);
      double overshot_depth = f.depth[cur] - f.depth[nx];
      double Se = lg_Se_from_theta(f.theta[cur], sp);
      f.psi[cur] = lg_h_from_Se(Se, sp);
      f.K[cur] = lg_K_from_Se(Se, sp, sp.Ksat * c.ff[layer_num]);  // lgar.cxx:1997,2023
      double theta_new = lg_theta_from_h(f.psi[cur], spn);
      double mbal_correction = overshot_depth * (current_theta - f.theta[nx]);
      double mbal_Z = mbal_correction / (theta_new - f.theta[nn]);
      if (isinf(mbal_Z)) mbal_Z = cum[layer_num + 1] / (1e3);
      double depth_new = cum[layer_num] + mbal_Z;
      f.depth[cur] = cum[layer_num];
      f.theta[nx] = theta_new;
      f.psi[nx] = f.psi[cur];
      f.depth[nx] = depth_new;
      f.layer[nx] = (signed char)(layer_num + 1);
      f.dzdt[nx] = f.dzdt[cur];
      f.dzdt[cur] = 0;
      f.tb[cur] = 1;
      f.tb[nx] = 0;
      cross_necessary = true;
      front_for_cross = nx + 1;
      if (depth_new > cum[num_layers] && f.layer[nx] == num_layers) theta_correction_necessary = true;
    }
    cur = cur + 1;
  }

  if (!cross_necessary) return;
  for (int wf = f.n - 1; wf != 0; wf--) {  // :2058-2084
    int i = wf - 1;
    const LgarLayer& sk = cls.lay[f.layer[i]];
    if (f.tb[i]) {
      f.psi[i] = f.psi[wf];
      f.theta[i] = lg_theta_from_h(f.psi[i], sk);
    }
    double Se = lg_Se_from_theta(f.theta[i], sk);
    f.K[i] = lg_K_from_Se(Se, sk, sk.Ksat);  // un-frozen Ks in the reference (lgar.cxx:2071,2080, quirk Q11)
  }
  if (!theta_correction_necessary) return;
  for (int wf = f.n - 1; wf != 0; wf--) {  // :2090-2159
    int i = wf - 1;
    if (i >= f.n) {
      c.status |= LGAR_FAIL_LIST;
      return;
    }
    double prior_mass = lg_mass_bal(c);
    if (f.depth[i] > cum[num_layers] && f.layer[i] == num_layers && wf == front_for_cross) {
      LG_PATH(c, LGAR_PATH_CROSS_LAYER_BOTTOM);
      f.depth[i] = cum[num_layers] + 1.E-6;
      lg_theta_mass_balance_correction(c, wf, prior_mass);
      double current_mass = lg_mass_bal(c);
      if (prior_mass > (current_mass + 100. * LG_MBAL_TOL)) {
        double max_increment = 1. / LG_TRUNCATION_DEPTH;
        double max_depth = cum[num_layers] + 1. / LG_TRUNCATION_DEPTH;
        double increment = 0.001;
        bool found_upper_bound = false;
        int iter_one_direction = 0;
        while (prior_mass > current_mass && increment < max_increment && f.depth[i] < max_depth) {
          iter_one_direction++;
          if (iter_one_direction > LG_MAX_ITER_MBAL) break;
          f.depth[i] += increment;
          current_mass = lg_mass_bal(c);
          if (current_mass >= prior_mass) {
            found_upper_bound = true;
            break;
          }
          increment *= 2.0;
        }
        if (found_upper_bound) {
          double low = f.depth[i] - increment;
          double high = f.depth[i];
          double cm2 = lg_mass_bal(c);
          while (fabs(cm2 - prior_mass) > LG_MBAL_TOL) {
            iter_one_direction++;
            double mid = 0.5 * (low + high);
            f.depth[i] = mid;
            cm2 = lg_mass_bal(c);
            if (cm2 >= prior_mass)
              high = mid;
            else
              low = mid;
            if (iter_one_direction > LG_MAX_ITER_MBAL) break;
          }
        }
      }
    }
  }
}

// lgar_wetting_front_cross_domain_boundary, lgar.cxx:2176-2251
LG_FN_NOINLINE double lg_cross_domain_boundary(LgCtx& c) {
  LgFronts& f = *c.f;
  const LgarColumnClass& cls = *c.cls;
  double domain_depth = cls.cum[cls.num_layers];
  double bottom_flux = 0.0;
  int length = f.n, cur = 0;
  for (int wf = 1; wf != length; wf++) {
    if (wf >= f.n) break;
    double flux_temp = 0.0;
    int nx = cur + 1, nn = cur + 2;
    if (nx >= f.n) {
      c.status |= LGAR_FAIL_LIST;
      return bottom_flux;
    }
    if (nn >= f.n && f.depth[cur] >= domain_depth) {
      const LgarLayer& sp = cls.lay[f.layer[cur]];
      flux_temp = (f.theta[cur] - f.theta[nx]) * (f.depth[cur] - f.depth[nx]);
      f.theta[nx] = f.theta[cur];
      double Se_k = lg_Se_from_theta(f.theta[cur], sp);
      f.psi[nx] = lg_h_from_Se(Se_k, sp);
      f.K[nx] = lg_K_from_Se(Se_k, sp, sp.Ksat * c.ff[f.layer[cur]]);  // lgar.cxx:2225-2230
      lg_front_delete(c, cur);
      bottom_flux += flux_temp;
      break;
    }
    cur = nx;
    if (flux_temp != 0.0) break;
  }
  return bottom_flux;
}

// lgar_check_dry_over_wet_wetting_fronts, lgar.cxx:2314-2338
LG_FN bool lg_check_dry_over_wet(const LgFronts& f) {
  for (int i = 0; i + 1 < f.n; i++)
    if ((f.theta[i] <= f.theta[i + 1]) && (f.layer[i] == f.layer[i + 1])) return true;
  return false;
}

// lgar_fix_dry_over_wet_wetting_fronts, lgar.cxx:2260-2307
LG_FN_NOINLINE void lg_fix_dry_over_wet(LgCtx& c, double* mass_change) {
  LgFronts& f = *c.f;
  int cur = 0, nx = (f.n > 1) ? 1 : -1;
  for (int l = 1; l <= f.n; l++) {
    if (nx >= 0) {
      if ((f.theta[cur] <= f.theta[nx]) && (f.layer[cur] == f.layer[nx])) {
        double prior_mass = lg_mass_bal(c);
        cur = lg_front_delete(c, cur);
        lg_theta_mass_balance_correction(c, cur + 1, prior_mass);
        double mass_after = lg_mass_bal(c);
        *mass_change += (mass_after - prior_mass);
      }
      cur = cur + 1;
      if (cur >= f.n)
        nx = -1;
      else
        nx = (cur + 1 < f.n) ? cur + 1 : -1;
    }
  }
}

// lgarto_correction_type_surf, lgar.cxx:3116-3169
LG_FN_NOINLINE int lg_correction_type(const LgCtx& c) {
  const LgFronts& f = *c.f;
  const double* cum = c.cls->cum;
  const int num_layers = c.cls->num_layers;
  const bool dry_over_wet = lg_check_dry_over_wet(f);  // loop-invariant in the reference (:3153)
  int cur = 0;
  for (int wf = 1; wf != f.n; wf++) {
    int nx = cur + 1;
    bool nn_null = (cur + 2 >= f.n);
    if (nx < f.n) {
      int la = f.layer[cur];
      if ((f.theta[cur] > f.theta[nx]) && (f.depth[cur] > f.depth[nx]) && (la == f.layer[nx]) && (!f.tb[nx])) return 1;
      if ((f.depth[cur] > cum[la]) && (f.depth[nx] == cum[la] && f.tb[nx]) && (f.theta[cur] > f.theta[nx]) && (la != num_layers))
        return 2;
    }
    if (nn_null && (f.depth[cur] > cum[f.layer[cur]])) return 3;
    if (dry_over_wet) return 4;
    cur = nx;
  }
  return 0;
}

// lgar_clean_redundant_fronts, lgar.cxx:3171-3195
LG_FN_NOINLINE void lg_clean_redundant(LgCtx& c) {
  LgFronts& f = *c.f;
  int cur = 0;
  for (int wf = 1; wf != f.n; wf++) {
    int nx = cur + 1;
    if ((f.layer[cur] == f.layer[nx]) && (fabs(f.theta[cur] - f.theta[nx]) < LG_THRESHOLD_NO_MOISTURE_DIFF)) {
      LG_PATH(c, LGAR_PATH_CLEAN_REDUNDANT);
      lg_front_delete(c, cur);
      break;
    }
    cur = nx;
  }
}

// lgar_move_wetting_fronts, lgar.cxx:1269-1866
LG_FN_NOINLINE double lg_move_wetting_fronts(LgCtx& c, double timestep_h, double free_drainage_cm, double* volin_cm,
                                             int wf_fd_demand, double old_mass, double mass_corr_fd,
                                             double* AET_demand_cm, const LgOld& old) {
  LgFronts& f = *c.f;
  const LgarColumnClass& cls = *c.cls;
  const double* cum = cls.cum;
  const int num_layers = cls.num_layers;
  const double column_depth = cum[num_layers];
  const int nwf = f.n;
  const double precip_mass_to_add = *volin_cm;
  double bottom_flux = 0.0;
  double dth[LGAR_MAXL + 1], dz[LGAR_MAXL + 1];
  *volin_cm = 0.0;

  for (int wf = nwf; wf != 0; wf--) {
    const int ci = wf - 1, ni = wf;
    if (ci >= f.n || ci >= old.n) {
      c.status |= LGAR_FAIL_LIST;
      return bottom_flux;
    }
    const int layer_num = f.layer[ci];
    const LgarLayer& sp = cls.lay[layer_num];
    int layer_below;
    if (wf == nwf)
      layer_below = layer_num + 1;
    else {
      if (ni >= f.n) {
        c.status |= LGAR_FAIL_LIST;
        return bottom_flux;
      }
      layer_below = f.layer[ni];
    }
    double actual_ET_demand = *AET_demand_cm;

    if ((wf < nwf) && (layer_below != layer_num)) {  // :1379-1387
      f.theta[ci] = lg_theta_from_h(f.psi[ni], sp);
      f.psi[ci] = f.psi[ni];
    }

    if (wf == nwf && layer_below != layer_num && nwf == num_layers) {  // :1403-1480
      f.depth[ci] += f.dzdt[ci] * timestep_h;
      double psi_old = old.psi[ci], psi_cm = f.psi[ci];
      double prior_mass = (old.depth[ci] - cum[layer_num - 1]) * (old.theta[ci] - 0.0);
      double new_mass = (f.depth[ci] - cum[layer_num - 1]) * (f.theta[ci] - 0.0);
      for (int k = 1; k < layer_num; k++) {
        const LgarLayer& sk = cls.lay[k];
        double theta_old = lg_theta_from_h(psi_old, sk);
        prior_mass += (sk.thick * theta_old);
        double theta = lg_theta_from_h(psi_cm, sk);
        new_mass += sk.thick * theta;
        dth[k] = 0.0;
        dz[k] = sk.thick;
      }
      dth[layer_num] = 0.0;
      dz[layer_num] = f.depth[ci] - cum[layer_num - 1];
      if (wf_fd_demand == wf) prior_mass += precip_mass_to_add - (free_drainage_cm + mass_corr_fd + actual_ET_demand);
      double theta_new = lg_theta_mass_balance(c, layer_num, psi_cm, new_mass, prior_mass, precip_mass_to_add, AET_demand_cm, dth, dz);
      actual_ET_demand = *AET_demand_cm;
      f.theta[ci] = fmax(sp.theta_r, fmin(theta_new, sp.theta_e));
      f.psi[ci] = lg_h_from_Se(lg_Se_from_theta(f.theta[ci], sp), sp);
    }

    if ((wf < nwf) && (layer_num == layer_below)) {  // :1489-1637
      if (ni >= old.n) {
        c.status |= LGAR_FAIL_LIST;
        return bottom_flux;
      }
      if (layer_num == 1) {
        double prior_mass = old.depth[ci] * (old.theta[ci] - old.theta[ni]);
        if (wf_fd_demand == wf) prior_mass += precip_mass_to_add - (free_drainage_cm + mass_corr_fd + actual_ET_demand);
        f.depth[ci] += f.dzdt[ci] * timestep_h;
        if (f.depth[ci] > column_depth) f.depth[ci] = column_depth + LG_TRUNCATION_DEPTH;
        if (f.dzdt[ci] == 0.0 && !f.tb[ci]) {
          // a front that was just created keeps its theta
        } else {
          double th = prior_mass / f.depth[ci] + f.theta[ni];
          if (th < sp.theta_r) {  // :1519-1537
            LG_PATH(c, LGAR_PATH_THETA_R_DELETE);
            double mass_before = lg_mass_bal(c) - f.depth[ci] * (f.theta[ci] - th);
            lg_front_delete(c, ci);
            double mass_after = lg_mass_bal(c);
            *AET_demand_cm = *AET_demand_cm - fabs(mass_before - mass_after);
            actual_ET_demand = *AET_demand_cm;
          } else {
            f.theta[ci] = fmax(sp.theta_r, fmin(sp.theta_e, th));
          }
        }
      } else {
        f.depth[ci] += f.dzdt[ci] * timestep_h;
        if (f.depth[ci] > column_depth) f.depth[ci] = column_depth + LG_TRUNCATION_DEPTH;
        double psi_old = old.psi[ci], psi_below_old = old.psi[ni];
        double psi_cm = f.psi[ci], psi_below = f.psi[ni];
        double prior_mass = (old.depth[ci] - cum[layer_num - 1]) * (old.theta[ci] - old.theta[ni]);
        double new_mass = (f.depth[ci] - cum[layer_num - 1]) * (f.theta[ci] - f.theta[ni]);
        for (int k = 1; k < layer_num; k++) {
          const LgarLayer& sk = cls.lay[k];
          double theta_old = lg_theta_from_h(psi_old, sk);
          double theta_below_old = lg_theta_from_h(psi_below_old, sk);
          prior_mass += (sk.thick * (theta_old - theta_below_old));
          double theta = lg_theta_from_h(psi_cm, sk);
          double theta_below = lg_theta_from_h(psi_below, sk);
          new_mass += sk.thick * (theta - theta_below);
          dth[k] = theta_below;
          dz[k] = sk.thick;
        }
        dth[layer_num] = f.theta[ni];
        dz[layer_num] = f.depth[ci] - cum[layer_num - 1];
        if (wf_fd_demand == wf) prior_mass += precip_mass_to_add - (free_drainage_cm + mass_corr_fd + actual_ET_demand);
        double theta_new = lg_theta_mass_balance(c, layer_num, psi_cm, new_mass, prior_mass, precip_mass_to_add, AET_demand_cm, dth, dz);
        actual_ET_demand = *AET_demand_cm;
        f.theta[ci] = fmax(sp.theta_r, fmin(theta_new, sp.theta_e));
      }
      // :1634-1635 (after a deletion slot ci holds the former `next`, same layer => same soil)
      f.psi[ci] = lg_h_from_Se(lg_Se_from_theta(f.theta[ci], sp), sp);
    }

    if (wf == 1) {  // :1644-1754
      wf_fd_demand = lg_wf_free_drainage(f);
      const int fdi = wf_fd_demand - 1;
      const double theta_e_k1 = cls.lay[f.layer[fdi]].theta_e;
      double mass_timestep = (old_mass + precip_mass_to_add) - (actual_ET_demand + free_drainage_cm + mass_corr_fd);
      if (!(old_mass > 0.0)) c.status |= LGAR_FAIL_OLD_MASS;
      if (fabs(f.theta[fdi] - theta_e_k1) < 1E-15) {
        double current_mass = lg_mass_bal(c);
        double mass_balance_error = fabs(current_mass - mass_timestep);
        double factor = 1.0;
        if (fdi + 1 < f.n) {
          if (fabs(f.theta[fdi] - f.theta[fdi + 1]) < 0.01) factor = factor / fabs(f.theta[fdi] - f.theta[fdi + 1]);
        }
        bool switched = false, aug = false, break_flag = false;
        double depth_new = f.depth[fdi];
        int iter = 0;
        const int speedup_thresh = LG_MAX_ITER_SAT / 10;
        if (fabs(mass_balance_error) > LG_MBAL_TOL) LG_PATH(c, LGAR_PATH_SATURATED_DEPTH);
        while (fabs(mass_balance_error) > LG_MBAL_TOL) {
          iter++;
          if (iter > LG_MAX_ITER_SAT) {
            break_flag = true;
            *AET_demand_cm = *AET_demand_cm + mass_balance_error;
            actual_ET_demand = *AET_demand_cm;
            break;
          }
          if ((iter > speedup_thresh) && (!aug)) {
            factor = factor * 100;
            aug = true;
          }
          if (current_mass < mass_timestep) {
            depth_new += 0.01 * factor;
            switched = false;
          } else {
            if (!switched) {
              switched = true;
              factor = factor * 0.1;
            }
            depth_new -= 0.01 * factor;
          }
          if (f.tb[fdi] && (f.layer[fdi] == num_layers)) depth_new = cum[num_layers];
          f.depth[fdi] = depth_new;
          current_mass = lg_mass_bal(c);
          mass_balance_error = fabs(current_mass - mass_timestep);
        }
        if (depth_new < LG_TRUNCATION_DEPTH) {
          depth_new = LG_TRUNCATION_DEPTH;
          f.depth[fdi] = depth_new;
        }
        if (isinf(f.depth[fdi])) {
          f.depth[fdi] = cum[num_layers] + 1.E-6;
          break_flag = true;
        }
        if (break_flag) {
          current_mass = lg_mass_bal(c);
          mass_timestep = (old_mass + precip_mass_to_add) - (actual_ET_demand + free_drainage_cm + mass_corr_fd);
          bottom_flux += mass_timestep - current_mass;
        }
      }
    }
  }

  // merge / cross / fix until nothing is left to correct, :1777-1814
  int correction = lg_correction_type(c);
  while (correction != 0) {
    if (c.status & (LGAR_FAIL_LIST | LGAR_FAIL_OVERFLOW)) break;
    LG_PATH(c, correction == 1 ? LGAR_PATH_MERGE : correction == 2 ? LGAR_PATH_CROSS_LAYER : correction == 3 ? LGAR_PATH_CROSS_BOTTOM : LGAR_PATH_DRY_OVER_WET);
    if (correction == 1) lg_merge(c);
    if (correction == 2) {
      double mass_before = lg_mass_bal(c);
      lg_cross_layer_boundary(c);
      double mass_after = lg_mass_bal(c);
      if (fabs(mass_before - mass_after) > 100. * LG_MBAL_TOL) *AET_demand_cm = *AET_demand_cm + (mass_before - mass_after);
    }
    if (correction == 3) {
      bottom_flux += lg_cross_domain_boundary(c);
      if (isnan(bottom_flux)) bottom_flux = 0.0;
    }
    if (correction == 4) {
      double mass_change = 0.0;
      lg_fix_dry_over_wet(c, &mass_change);
      *AET_demand_cm = *AET_demand_cm - mass_change;
    }
    correction = lg_correction_type(c);
  }

  for (int i = 0; i < f.n; i++) {  // :1824-1847
    const LgarLayer& sk = cls.lay[f.layer[i]];
    double Se = lg_Se_from_theta(f.theta[i], sk);
    if (f.psi[i] > 1.0) f.psi[i] = lg_h_from_Se(Se, sk);
    f.K[i] = lg_K_from_Se(Se, sk, c.ff[f.layer[i]] * sk.Ksat);  // :1836
  }
  if (f.n == 1) {  // :1858-1863
    if (f.depth[0] != cum[1]) f.depth[0] = cum[1];
  }
  return bottom_flux;
}

// lgar_insert_water, lgar.cxx:2348-2484 (precip argument is a rate, quirk Q4)
LG_FN_NOINLINE double lg_insert_water(LgCtx& c, double timestep_h, double AET_demand_cm, double free_drainage_cm,
                                      double* ponded_depth_cm, double* volin, double precip_rate, int wf_fd_demand) {
  const LgFronts& f = *c.f;
  const LgarColumnClass& cls = *c.cls;
  const LgarConfig& cfg = *c.cfg;
  const double* cum = cls.cum;
  double f_p, runoff = 0.0;
  double h_p = fmax(*ponded_depth_cm - precip_rate * timestep_h, 0.0);
  const int fdi = wf_fd_demand - 1;
  const int layer_fp = f.layer[fdi];
  const LgarLayer& sp = cls.lay[layer_fp];
  // Ks of the infiltrating layer times the frozen factor of the layer of the TOP front (lgar.cxx:2384,2401, quirk Q4)
  const double Ksat = sp.Ksat * c.ff[f.layer[0]];
  double Geff;
  if (f.n == cls.num_layers)
    Geff = 0.0;
  else
    Geff = lg_Geff(cfg, f.theta[fdi + 1], sp.theta_e, sp, Ksat);
  if (layer_fp == 1) {
    f_p = Ksat * (1 + (Geff + h_p) / f.depth[fdi]);
  } else {
    double bottom_sum = (f.depth[fdi] - cum[layer_fp - 1]) / Ksat;
    for (int k = 1; k < layer_fp; k++)
      bottom_sum += (cum[layer_fp - k] - cum[layer_fp - (k + 1)]) / (cls.lay[layer_fp - k].Ksat * c.ff[layer_fp - k]);  // :2420
    f_p = (f.depth[fdi] / bottom_sum) + ((Geff + h_p) * Ksat / (f.depth[fdi]));
  }
  double current_mass = lg_mass_bal(c);
  double cap = (cls.max_storage + free_drainage_cm + AET_demand_cm - current_mass) / timestep_h;
  if (f_p > cap) f_p = cap;
  double ponded_temp = fmax(*ponded_depth_cm - f_p * timestep_h - 0.0, 0.0);
  double fp_cm = f_p * timestep_h + 0.0 / timestep_h;
  const double ponded_depth_max_cm = cls.par.ponded_depth_max_cm;
  if (ponded_depth_max_cm > 0.0) {
    if (ponded_temp < ponded_depth_max_cm) {
      *volin = fmin(*ponded_depth_cm, fp_cm);
      *ponded_depth_cm = *ponded_depth_cm - *volin;
      return 0.0;
    } else if (ponded_temp > ponded_depth_max_cm) {
      runoff = ponded_temp - ponded_depth_max_cm;
      *ponded_depth_cm = ponded_depth_max_cm;
      *volin = fp_cm;
      return runoff;
    }
    // equality: the reference falls through without touching anything
  } else {
    *volin = fmin(*ponded_depth_cm, fp_cm);
    runoff = *ponded_depth_cm < fp_cm ? 0.0 : (*ponded_depth_cm - *volin);
    *ponded_depth_cm = 0.0;
  }
  return runoff;
}

// lgar_calc_dry_depth, lgar.cxx:2577-2626
LG_FN_NOINLINE double lg_calc_dry_depth(LgCtx& c, double timestep_h) {
  const LgFronts& f = *c.f;
  const int layer_num = f.layer[0];
  const LgarLayer& sp = c.cls->lay[layer_num];
  const double Ksat = sp.Ksat * c.ff[layer_num];  // lgar.cxx:2602
  double tau = timestep_h * Ksat / (sp.theta_e - f.theta[0]);
  double Geff = lg_Geff(*c.cfg, f.theta[0], sp.theta_e, sp, Ksat);
  double dry_depth = 0.5 * (tau + sqrt(tau * tau + 4.0 * tau * Geff));
  return fmin(c.cls->cum[layer_num], dry_depth);
}

// lgar_create_surficial_front, lgar.cxx:2490-2568
LG_FN_NOINLINE void lg_create_surficial_front(LgCtx& c, double* ponded_depth_cm, double* volin, double dry_depth, double theta1) {
  LG_PATH(c, LGAR_PATH_SURFICIAL);
  LgFronts& f = *c.f;
  const LgarLayer& sp = c.cls->lay[1];
  double delta_theta = sp.theta_e - theta1;
  double theta_new;
  if (dry_depth * delta_theta > (*ponded_depth_cm)) {
    *volin = *ponded_depth_cm;
    theta_new = fmin(theta1 + (*ponded_depth_cm) / dry_depth, sp.theta_e);
    lg_front_insert_first(c, dry_depth, theta_new, 1, 0);
    *ponded_depth_cm = 0.0;
  } else {
    *volin = dry_depth * delta_theta;
    *ponded_depth_cm -= dry_depth * delta_theta;
    theta_new = sp.theta_e;
    lg_front_insert_first(c, dry_depth, sp.theta_e, 1, 0);
  }
  if (c.status & LGAR_FAIL_OVERFLOW) return;
  double Se = lg_Se_from_theta(theta_new, sp);
  f.psi[0] = lg_h_from_Se(Se, sp);
  f.K[0] = lg_K_from_Se(Se, sp, sp.Ksat * c.ff[1]) * c.ff[1];  // the factor is applied twice in the reference (lgar.cxx:2542,2547, quirk Q11)
  f.dzdt[0] = 0.0;
  if (f.n > 1) {
    bool had_to_merge = false;
    while ((f.n > 1) && (f.depth[0] > f.depth[1]) && (f.layer[0] == f.layer[1]) && !f.tb[1]) {
      lg_merge(c);
      had_to_merge = true;
      if (c.status & LGAR_FAIL_LIST) break;
    }
    if (had_to_merge) lg_cross_layer_boundary(c);
  }
}

// lgar_dzdt_calc, lgar.cxx:2764-2945
LG_FN_NOINLINE void lg_dzdt_calc(LgCtx& c, double h_p, double subtimestep_h, bool switch_caching, int cache_count, int new_front) {
  LgFronts& f = *c.f;
  const LgarColumnClass& cls = *c.cls;
  const double* cum = cls.cum;
  int count_correct_for_crossing = 0;
  for (int i = 0; i < f.n; i++) {
    // Only the fronts that are not to-bottom (about one in four) need the Green-Ampt evaluation below.  The cheap
    // cases are skipped in this inner loop, so that the k-th trip of the outer loop is the k-th MOVING front of
    // every lane, whatever its index: the lanes of a warp run the expensive part together (it ran at 6.5 of 32
    // lanes when the trips were aligned by front index).  Same visits in the same order per column.
    while (i + 1 < f.n && f.tb[i] && !(f.K[i] < 0)) {
      f.dzdt[i] = 0.0;
      i++;
    }
    double dzdt = 0.0;
    const int layer_num = f.layer[i];
    const double K_cm_per_h = f.K[i];
    const LgarLayer& sp = cls.lay[layer_num];
    if (K_cm_per_h < 0) {
      c.status |= LGAR_FAIL_K_NEGATIVE;
      return;
    }
    if (i + 1 >= f.n) break;
    double theta1 = f.theta[i + 1], theta2 = f.theta[i];
    double bottom_sum = 0.0;
    if (f.tb[i]) {
      f.dzdt[i] = 0.0;
      continue;
    } else if (layer_num > 1) {
      bottom_sum += (f.depth[i] - cum[layer_num - 1]) / K_cm_per_h;
    }
    if (theta1 > theta2) {
      c.status |= LGAR_FAIL_THETA_ORDER;
      return;
    }
    const double Ksat = sp.Ksat * c.ff[layer_num];  // lgar.cxx:2824
    double Geff = lg_Geff(*c.cfg, theta1, theta2, sp, Ksat);
    double delta_theta = f.theta[i] - f.theta[i + 1];
    if (layer_num == 1) {
      if (delta_theta > 0)
        dzdt = 1.0 / delta_theta * (Ksat * (Geff + h_p) / f.depth[i] + f.K[i]);
      else
        dzdt = 0.0;
    } else {
      double denominator = bottom_sum;
      for (int k = 1; k < layer_num; k++) {  // quirk Q3: K of layer (layer_num-k) with the thickness of layer k
        const LgarLayer& sl = cls.lay[layer_num - k];
        double theta_prev = lg_theta_from_h(f.psi[i], sl);
        double K_prev = lg_K_from_Se(lg_Se_from_theta(theta_prev, sl), sl, sl.Ksat * c.ff[layer_num - k]);  // :2877
        denominator += (cum[k] - cum[k - 1]) / K_prev;
      }
      if (delta_theta > 0)
        dzdt = (1.0 / delta_theta) * ((f.depth[i] / denominator) + Ksat * (Geff + h_p) / f.depth[i]);
      else
        dzdt = 0.0;
    }
    if ((dzdt == 0.0) && !f.tb[i]) dzdt = 1e-9;
    if (dzdt > 1e4) dzdt = 1e4;
    if (dzdt > 100 * cls.largest_Ks) dzdt = 100 * cls.largest_Ks;
    if (dzdt < -100 * cls.largest_Ks) dzdt = -100 * cls.largest_Ks;
    if (f.tb[i + 1] && dzdt * subtimestep_h + f.depth[i] > LG_FACTOR_LIMITS_CROSSING * (cum[layer_num])) {
      count_correct_for_crossing++;
      dzdt = fabs(f.depth[i] - LG_FACTOR_LIMITS_CROSSING * (cum[layer_num])) / subtimestep_h +
             count_correct_for_crossing * LG_DEPTH_AVOIDS_SAME;
    }
    if (switch_caching) {
      if ((i + 1) != new_front) dzdt = dzdt * (cache_count);
    }
    f.dzdt[i] = dzdt;
  }
}

// calc_aet, src/aet.cxx:17-77.  psi_50 (the reference's psi_wp_cm) is a class constant.
LG_FN double lg_calc_aet(const LgCtx& c, double PET_cm_per_h, double dt_h) {
  const LgFronts& f = *c.f;
  double aet = 0.0;
  if (f.psi[0] < 1.E6) {
    double r = f.psi[0] / c.cls->psi_50_cm[f.layer[0]];
    double h_ratio = 1.0 + lg_pow3(r);
    aet = PET_cm_per_h * (1 / h_ratio) * dt_h;
    if (aet < 0)
      aet = 0.0;
    else if (aet > (PET_cm_per_h * dt_h))
      aet = PET_cm_per_h * dt_h;
  }
  return aet;
}

// calc_CR_Q, src/conceptual_reservoir.cxx:7-45
LG_FN double lg_calc_CR_Q_inl(const LgarColumnParams& cfg, double dt, double in_cm_per_h, double* fast, double* slow) {
  double input_slow = in_cm_per_h * cfg.frac_slow;
  double input_fast = in_cm_per_h - input_slow;
  double Q_fast = 0.0;
  if (!(*fast < 0.01)) Q_fast = dt * (cfg.a * lg_pow(*fast, cfg.b));
  double delta_fast = dt * input_fast - Q_fast;
  if (*fast + delta_fast > 0.0) {
    *fast += delta_fast;
  } else {
    Q_fast = *fast + dt * input_fast;
    *fast = 0.0;
  }
  double Q_slow = 0.0;
  if (!(*slow < 0.01)) Q_slow = dt * (cfg.a_slow * lg_pow(*slow, cfg.b_slow));
  double delta_slow = dt * input_slow - Q_slow;
  if (*slow + delta_slow > 0.0) {
    *slow += delta_slow;
  } else {
    Q_slow = *slow + dt * input_slow;
    *slow = 0.0;
  }
  return Q_fast + Q_slow;
}
LG_FN_NOINLINE double lg_calc_CR_Q(const LgarColumnParams& cfg, double dt, double in_cm_per_h, double* fast, double* slow) { return lg_calc_CR_Q_inl(cfg, dt, in_cm_per_h, fast, slow); }

// ------------------------------------------------------------------------------------------------
// L3: BmiLGAR::Update(), reference src/bmi_lgar.cxx:133-895
// ------------------------------------------------------------------------------------------------

// What one forcing step hands to the epilogue (GIUH, outputs, cumulative sums), all in cm.
struct LgStepVols {
  double precip, PET, AET, volin, volrech, volrunoff, volQ_CR, volend, volCRend, local_mb;
};

struct LgPrologue {
  double subtimestep_h;
  int subcycles;
  bool cache_fluxes, switch_caching;
  int cache_count;
};

// adaptive sub-step + flux-cache decision, bmi_lgar.cxx:260-318.  Pure: commits nothing.
LG_FN LgPrologue lg_prologue_inl(const LgarConfig& cfg, const LgScalars& s, double precip_mm_h) {
  LgPrologue p;
  double timestep_h = s.timestep_h;
  p.subtimestep_h = timestep_h;
  if (cfg.adaptive_timestep) {
    double dt = cfg.forcing_resolution_h;
    if (precip_mm_h > 10.0 || s.volon > 0.0)
      dt = cfg.timestep_h;
    else if (precip_mm_h > 0.0)
      dt = cfg.timestep_h * 2.0;
    dt = fmin(dt, cfg.forcing_resolution_h);
    p.subtimestep_h = dt;
    timestep_h = dt;
  }
  const bool caching_at_start = s.cache_fluxes != 0;
  bool cache = caching_at_start;
  double wf_free_drain_dzdt = caching_at_start ? 0.0 : s.fd_dzdt;
  if (cfg.allow_flux_caching) {
    if (p.subtimestep_h == cfg.forcing_resolution_h) {
      if ((precip_mm_h < LG_PRECIP_THRESHOLD_MM_PER_H) && ((s.prev_aet / (s.prev_pet + LG_PET_EPSILON)) < LG_AET_PET_RATIO_THRESHOLD) &&
          (s.cache_count != LG_CACHE_RESET_STEPS) && (s.volon < LG_VOLON_THRESHOLD_CM) && (s.prev_rech < LG_BOTTOM_FLUX_THRESHOLD_CM))
        cache = true;
      if (wf_free_drain_dzdt > LG_THRESHOLD_DZDT) cache = false;
      if (s.volon >= LG_VOLON_THRESHOLD_CM) cache = false;
      if ((s.cache_count == LG_CACHE_RESET_STEPS) || (precip_mm_h >= LG_PRECIP_THRESHOLD_MM_PER_H)) cache = false;
    } else {
      cache = false;
    }
  }
  p.cache_fluxes = cache;
  p.switch_caching = caching_at_start && !cache;
  p.cache_count = s.cache_count + (cache ? 1 : 0);
  p.subcycles = (int)(cfg.forcing_resolution_h / timestep_h + 1.0e-08);
  return p;
}
LG_FN_NOINLINE LgPrologue lg_prologue(const LgarConfig& cfg, const LgScalars& s, double precip_mm_h) { return lg_prologue_inl(cfg, s, precip_mm_h); }

// One sub-cycle's tail shared by the cached and the computed branch: reservoir, preferential-flow
// flag, local mass balance (bmi_lgar.cxx:697-737).
struct LgCycle {
  double precip_rate, precip_for_CR, ponded_flux_for_CR, precip_sub, PET_sub;
  double volstart, AET, volin, volon_sub, runoff, volrech, fd_for_CR, volon_ts;
  bool top_near_sat;
};

LG_FN double lg_cycle_tail_inl(const LgarConfig& cfg, LgScalars& s, LgCycle& cy, double dt, double volend, LgStepVols& v) {
  s.precip_prev = cy.precip_sub;
  double volCRstart = s.cr_fast + s.cr_slow;
  double volQ_CR = lg_calc_CR_Q_inl(*s.rp, dt, cy.precip_for_CR + cy.ponded_flux_for_CR + cy.fd_for_CR / dt, &s.cr_fast, &s.cr_slow);
  v.volQ_CR += volQ_CR;
  double volCRend = s.cr_fast + s.cr_slow;
  v.volCRend = volCRend;
  s.runoff_prev = ((cy.runoff > LG_SMALL_EPS) || cy.top_near_sat) ? 1 : 0;
  cy.precip_sub += cy.precip_for_CR * dt;
  double local_mb = cy.volstart + cy.precip_sub + cy.volon_ts - cy.runoff - volQ_CR - volCRend + volCRstart - cy.AET -
                    cy.volon_sub - cy.volrech - volend;
  v.volin += cy.volin;
  v.volrech += cy.volrech;
  v.volrunoff += cy.runoff;
  v.AET += cy.AET;
  v.volend = volend;
  v.local_mb = local_mb;
  return local_mb;
}
LG_FN_NOINLINE double lg_cycle_tail(const LgarConfig& cfg, LgScalars& s, LgCycle& cy, double dt, double volend, LgStepVols& v) { return lg_cycle_tail_inl(cfg, s, cy, dt, volend, v); }

// start of a sub-cycle: preferential flow split and forcing amounts, bmi_lgar.cxx:362-416
LG_FN void lg_cycle_head(const LgarConfig& cfg, const LgScalars& s, LgCycle& cy, double precip_mm_h, double pet_mm_h, double dt,
                         double& volon_ts, double& PET_rate, double& ponded_depth) {
  cy.precip_for_CR = 0.0;
  cy.ponded_flux_for_CR = 0.0;
  cy.precip_rate = precip_mm_h * 0.1;
  if (s.runoff_prev) {
    double total = cy.precip_rate;
    const double frac_to_CR = s.rp->frac_to_CR;
    cy.precip_for_CR = frac_to_CR * total;
    cy.precip_rate = (1.0 - frac_to_CR) * total;
    cy.ponded_flux_for_CR = volon_ts / dt * frac_to_CR;
    volon_ts = volon_ts * (1 - frac_to_CR);
  }
  cy.AET = 0.0;
  cy.volin = 0.0;
  cy.volon_sub = 0.0;
  cy.runoff = 0.0;
  cy.volrech = 0.0;
  cy.fd_for_CR = 0.0;
  cy.top_near_sat = false;
  PET_rate = pet_mm_h * 0.1;
  ponded_depth = cy.precip_rate * dt;
  ponded_depth += volon_ts;
  cy.precip_sub = cy.precip_rate * dt;
  cy.PET_sub = PET_rate * dt;
}

// A forcing step that replays cached fluxes (bmi_lgar.cxx:683-695): one sub-cycle, fronts untouched.
LG_FN void lg_step_cached_inl(const LgarConfig& cfg, LgScalars& s, const LgPrologue& pro, double precip_mm_h, double pet_mm_h, LgStepVols& v) {
  const double dt = pro.subtimestep_h;
  double volon_ts = s.volon, PET_rate, ponded;
  LgCycle cy;
  lg_cycle_head(cfg, s, cy, precip_mm_h, pet_mm_h, dt, volon_ts, PET_rate, ponded);
  cy.volstart = s.storage;
  v.PET += fmax(cy.PET_sub, 0.0);
  s.acc_pet += v.PET;
  if (!cfg.free_drainage_to_CR)
    cy.volrech += s.prev_rech;
  else
    cy.fd_for_CR += s.prev_rech;
  s.acc_fd += s.prev_rech;
  cy.volon_ts = volon_ts;
  lg_cycle_tail_inl(cfg, s, cy, dt, s.storage, v);
  s.volon = cy.volon_sub;
  s.prev_aet = cy.AET;
  s.prev_pet = cy.PET_sub;
}
LG_FN_NOINLINE void lg_step_cached(const LgarConfig& cfg, LgScalars& s, const LgPrologue& pro, double precip_mm_h, double pet_mm_h, LgStepVols& v) { lg_step_cached_inl(cfg, s, pro, precip_mm_h, pet_mm_h, v); }

// A forcing step that computes fluxes (bmi_lgar.cxx:356-805 with cache_fluxes == false).
LG_FN_NOINLINE void lg_step_active(const LgarConfig& cfg, const LgarColumnClass& cls, LgScalars& s, LgFronts& f, const LgPrologue& pro,
                                   double precip_mm_h, double pet_mm_h, LgStepVols& v, const double* frozen_factor, unsigned long long* paths = nullptr, bool use_ff = false) {
  LgCtx c;
  c.paths = paths;
  c.use_ff = use_ff;
  c.cfg = &cfg;
  c.cls = &cls;
  c.f = &f;
  c.status = 0;
  c.ff = frozen_factor;
  const double dt = pro.subtimestep_h;
  const int num_layers = cls.num_layers;
  double volon_ts = s.volon;
  double volend_sub = s.storage;  // == lgar_calc_mass_bal(head) at entry (bmi_lgar.cxx:202,220)
  double last_AET = 0.0, last_PET_sub = 0.0, last_volrech = 0.0;
  LgOld old;

  for (int cycle = 1; cycle <= pro.subcycles; cycle++) {
    LG_PATH(c, LGAR_PATH_ACTIVE_STEP);
    LgCycle cy;
    double PET_rate, ponded;
    lg_cycle_head(cfg, s, cy, precip_mm_h, pet_mm_h, dt, volon_ts, PET_rate, ponded);
    old.n = f.n;
    for (int i = 0; i < f.n; i++) {  // state_previous = listCopy(head), :370-374
      old.depth[i] = f.depth[i];
      old.theta[i] = f.theta[i];
      old.psi[i] = f.psi[i];
    }
    cy.volstart = lg_mass_bal(c);
    const double mass_check = cy.volstart;  // :422 (same list, same sum)
    const int wf_fd = lg_wf_free_drainage(f);
    const double min_water_fd = lg_min_water_fd(c, wf_fd);
    const double storage_fd = lg_storage_fd(c, wf_fd);
    double mass_corr_fd = 0.0, free_drainage = 0.0, temp_rch = 0.0;

    if (pro.switch_caching) {  // :436-441 (quirk Q6)
      PET_rate += s.acc_pet;
      s.acc_pet = 0.0;
      mass_corr_fd += s.acc_fd;
      s.acc_fd = 0.0;
    }
    if (cfg.free_drainage_enabled) {  // :443-461
      const int last = f.n - 1;
      free_drainage += dt * f.K[last];
      int it = 0;
      while ((mass_check - free_drainage < cls.min_storage) || (storage_fd - free_drainage < min_water_fd)) {
        free_drainage *= 0.5;
        if (it > 5) {
          free_drainage = 0.0;
          break;
        }
        it++;
      }
      if (free_drainage < 1.E-7) free_drainage = 0.0;
      if (f.psi[last] > 1.E6) free_drainage = 0.0;
    }
    const double precip_prev_sub = s.precip_prev;
    double AET = 0.0;
    if (PET_rate > 0.0) AET = lg_calc_aet(c, PET_rate, dt);
    {
      int it = 0;  // :487-495
      while ((mass_check - AET < cls.min_storage) || (storage_fd - AET < min_water_fd)) {
        AET *= 0.5;
        if (it > 5) {
          AET = 0.0;
          break;
        }
        it++;
      }
      it = 0;  // :497-510
      while (((mass_check - AET - free_drainage - mass_corr_fd) < cls.min_storage) ||
             ((storage_fd - AET - free_drainage - mass_corr_fd) < min_water_fd)) {
        AET *= 0.5;
        free_drainage *= 0.5;
        mass_corr_fd *= 0.5;
        if (it > 5) {
          AET = 0.0;
          free_drainage = 0.0;
          mass_corr_fd = 0.0;
          break;
        }
        it++;
      }
    }
    v.precip += cy.precip_sub + cy.precip_for_CR * dt;
    v.PET += fmax(cy.PET_sub, 0.0);
    volon_ts = fmax(volon_ts, 0.0);
    volon_ts = volon_ts > LG_SMALL_EPS ? volon_ts : 0.0;

    // create a surficial front? :522-535 (quirk Q8)
    const double theta_e_top = cls.lay[f.layer[0]].theta_e;
    const bool top_saturated = (f.theta[0] + LG_SMALL_EPS) >= theta_e_top;
    cy.top_near_sat = f.theta[0] > theta_e_top * cls.par.spf_factor;
    bool create = (precip_prev_sub == 0.0 && cy.precip_sub > 0.0 && volon_ts == 0) ||
                  ((cy.precip_sub > 0.0 || volon_ts > 0) && (f.n == num_layers));
    if (top_saturated || volon_ts > 0.0) create = false;

    double volin = 0.0, runoff = 0.0, volon_sub = 0.0;
    if (create) {  // :547-595
      double temp_pd = 0.0;
      temp_rch = lg_move_wetting_fronts(c, dt, free_drainage, &temp_pd, wf_fd, volend_sub, mass_corr_fd, &AET, old);
      double dry_depth = lg_calc_dry_depth(c, dt);
      lg_create_surficial_front(c, &ponded, &volin, dry_depth, f.theta[0]);
    }
    if (ponded > 0 && !create) {  // :601-618
      runoff = lg_insert_water(c, dt, AET, free_drainage, &ponded, &volin, cy.precip_rate, wf_fd);
      volon_sub = ponded;
      if (runoff < 0) c.status |= LGAR_FAIL_NEG_RUNOFF;
    } else {  // :619-632
      if (ponded < cls.par.ponded_depth_max_cm) {
        volon_sub = ponded;
        ponded = 0.0;
        runoff = 0.0;
      } else {
        runoff = (ponded - cls.par.ponded_depth_max_cm);
        volon_sub = cls.par.ponded_depth_max_cm;
        ponded = cls.par.ponded_depth_max_cm;
      }
    }
    if (!create && !(c.status & LGAR_FAIL_MASK)) {  // :638-652
      double volin_tmp = volin;
      temp_rch = lg_move_wetting_fronts(c, dt, free_drainage, &volin_tmp, wf_fd, volend_sub, mass_corr_fd, &AET, old);
    }
    if (!(c.status & LGAR_FAIL_MASK)) {
      lg_clean_redundant(c);  // :654
      int new_front = (pro.switch_caching && create) ? 1 : -1;
      lg_dzdt_calc(c, ponded, dt, pro.switch_caching, s.cache_count, new_front);  // :664-666
    }
    if (pro.switch_caching) s.cache_count = 1;
    double volrech = temp_rch;  // :672-679
    if (!cfg.free_drainage_to_CR) {
      volrech += free_drainage;
    } else {
      cy.fd_for_CR += free_drainage + volrech;
      volrech = 0.0;
    }
    cy.AET = AET;
    cy.volin = volin;
    cy.runoff = runoff;
    cy.volon_sub = volon_sub;
    cy.volrech = volrech;
    cy.volon_ts = volon_ts;
    volend_sub = lg_mass_bal(c);
    double local_mb = lg_cycle_tail(cfg, s, cy, dt, volend_sub, v);
    volon_ts = volon_sub;
    last_AET = AET;
    last_PET_sub = cy.PET_sub;
    last_volrech = volrech;
    if (!(fabs(local_mb) <= cfg.mbal_tol)) c.status |= (isnan(local_mb) ? LGAR_FAIL_NAN : LGAR_FAIL_MBAL);  // :745-788
    if (!(f.depth[0] > 0.0)) c.status |= LGAR_FAIL_TOP_DEPTH;                                              // :795
    if (c.status & LGAR_FAIL_MASK) break;
  }

  int tbc = 0;  // :824-834
  for (int i = 0; i < f.n; i++) tbc += f.tb[i] ? 1 : 0;
  if (tbc > num_layers) c.status |= LGAR_FAIL_TO_BOTTOM;

  s.volon = volon_ts;
  s.prev_aet = last_AET;
  s.prev_pet = last_PET_sub;
  s.prev_rech = last_volrech;
  s.storage = volend_sub;
  s.fd_dzdt = f.dzdt[lg_wf_free_drainage(f) - 1];
  s.status |= c.status;
}

#endif  // LGAR_PHYSICS_CUH
