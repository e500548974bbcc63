// lgar_c_b200/csrc/lgar_types.h
//
// Plain-old-data shared by the host engine (lgar_engine.cu) and the device physics
// (lgar_physics.cuh).  The layout replaces the reference's heap objects:
//   struct wetting_front linked list (reference include/all.hxx:45-56, src/linked_list.cxx)
//       -> fixed-capacity structure-of-arrays in HBM, column index fastest varying
//   struct soil_properties_ + lgar_bmi_parameters (all.hxx:66-79,102-169)
//       -> one read-only ColumnClass row per distinct (layering, soil tuple), shared by columns
//   struct lgar_mass_balance_variables (all.hxx:172-219)
//       -> per-column scalar arrays
#ifndef LGAR_TYPES_H
#define LGAR_TYPES_H

#include <stdint.h>

#define LGAR_W 32        // front capacity per column.  The shipped configs stay below 15, but a full year of
                         // BASELINE config 5 reaches 19 fronts in a 10 000-column sample (tools/parity_sampling_full.py; the
                         // tail falls by ~2.3x per front, so ~27 over 10 M columns): 16 froze 0.2 % of the columns.
#define LGAR_FW (LGAR_W / 16)  // 64-bit flag words per column (16 nibbles each)
#define LGAR_MAXL 4      // soil layers per column (reference MAX_NUM_SOIL_LAYERS, all.hxx:38)
#define LGAR_MAXG 32     // GIUH ordinates after re-sampling (lgar.cxx:918-929)
#define LGAR_NOUT 12     // per-step scalar outputs (bmi_lgar.hxx:168-181)

// per-column status word.  Low 16 bits: the column hit a condition on which the reference calls
// abort()/exit()/assert; it is frozen from then on and its outputs are NaN.
#define LGAR_OK 0
#define LGAR_FAIL_MBAL 1           // bmi_lgar.cxx:785-788
#define LGAR_FAIL_NEG_RUNOFF 2     // bmi_lgar.cxx:614-617
#define LGAR_FAIL_K_NEGATIVE 4     // lgar.cxx:2805-2812
#define LGAR_FAIL_THETA_ORDER 8    // lgar.cxx:2850-2853
#define LGAR_FAIL_OVERFLOW 16      // more than LGAR_W fronts
#define LGAR_FAIL_LIST 32          // list walked off its end (reference: NULL dereference)
#define LGAR_FAIL_TOP_DEPTH 64     // bmi_lgar.cxx:795
#define LGAR_FAIL_TO_BOTTOM 128    // bmi_lgar.cxx:830-834
#define LGAR_FAIL_OLD_MASS 256     // lgar.cxx:1654
#define LGAR_FAIL_MERGE_DEPTH 512  // lgar.cxx:1916
#define LGAR_FAIL_NAN 1024         // non-finite mass balance
#define LGAR_FAIL_MASK 0xFFFF
#define LGAR_WARN_NEG_FORCING 0x10000

// Path counters (lgarb_set_path_counters): how often each rare branch of the reference ran -- the evidence that the tests reach
// them.  Same numbering as LGO_PATH_* of oracle/lgar_oracle.h, so the engine's counts can be compared with the oracle's.
enum {
  LGAR_PATH_MERGE = 0,          // lgar_merge_wetting_fronts, lgar.cxx:1875-1953 (correction type 1)
  LGAR_PATH_CROSS_LAYER,        // lgar_wetting_fronts_cross_layer_boundary, :1963-2167 (type 2)
  LGAR_PATH_CROSS_BOTTOM,       // lgar_wetting_front_cross_domain_boundary, :2176-2251 (type 3)
  LGAR_PATH_DRY_OVER_WET,       // lgar_fix_dry_over_wet_wetting_fronts, :2260-2307 (type 4)
  LGAR_PATH_THETA_R_DELETE,     // front deleted because theta < theta_r, :1519-1537
  LGAR_PATH_SATURATED_DEPTH,    // saturated front deepened until the mass balance closes, :1687-1723
  LGAR_PATH_CROSS_LAYER_BOTTOM, // layer crossing that also passes the domain bottom: depth search, :2090-2159
  LGAR_PATH_SURFICIAL,          // lgar_create_surficial_front, :2490-2568
  LGAR_PATH_CLEAN_REDUNDANT,    // lgar_clean_redundant_fronts deletes a front, :3171-3195
  LGAR_PATH_SEARCH,             // lgar_theta_mass_balance iterates, :2990
  LGAR_PATH_SEARCH_CONVERGED,   // ... and ends within the tolerance
  LGAR_PATH_SEARCH_STEP_EXIT,   // ... on |psi - psi_prev| < 1e-15 and factor < 1e-13, :3063
  LGAR_PATH_SEARCH_NOCHANGE,    // ... on five trips without mass change, :3070-3076
  LGAR_PATH_SEARCH_MAXITER,     // ... on the trip limit, :3078
  LGAR_PATH_SEARCH_PSI_UPPER,   // ... on psi > 1e7 after 1000 trips, :3081
  LGAR_PATH_SEARCH_PSI_ZERO,    // ... on psi <= 0 twice, :3084
  LGAR_PATH_CORR_SEARCH,        // lgar_theta_mass_balance_correction iterates, :3307
  LGAR_PATH_CACHED_STEP,        // sub-step that replays cached fluxes, bmi_lgar.cxx:683-695
  LGAR_PATH_ACTIVE_STEP,        // sub-step that computes fluxes
  LGAR_NPATH = 24
};

// order of the per-step scalar outputs, metres (bmi_lgar.cxx:882-893)
enum {
  LGAR_OUT_PRECIP = 0, LGAR_OUT_PET, LGAR_OUT_AET, LGAR_OUT_RUNOFF, LGAR_OUT_GIUH_RUNOFF,
  LGAR_OUT_STORAGE, LGAR_OUT_Q, LGAR_OUT_INFIL, LGAR_OUT_PERC, LGAR_OUT_Q_CR, LGAR_OUT_MBAL,
  LGAR_OUT_CR_STORAGE
};

// cumulative volumes (cm) kept for the global mass balance, lgar.cxx:1162-1185
enum {
  LGAR_CUM_PRECIP = 0, LGAR_CUM_IN, LGAR_CUM_AET, LGAR_CUM_RECH, LGAR_CUM_RUNOFF, LGAR_CUM_Q,
  LGAR_CUM_Q_CR, LGAR_CUM_PET, LGAR_CUM_GIUH, LGAR_CUM_RUNOFF_CR, LGAR_NCUM
};

// persistent per-column f64 scalars (SURVEY.md section 8 a-S)
enum {
  LGAR_S_VOLON = 0, LGAR_S_PRECIP_PREV, LGAR_S_CR_FAST, LGAR_S_CR_SLOW, LGAR_S_PREV_AET,
  LGAR_S_PREV_PET, LGAR_S_PREV_RECH, LGAR_S_ACC_PET, LGAR_S_ACC_FD, LGAR_S_TIMESTEP_H,
  LGAR_S_STORAGE,  // lgar_calc_mass_bal(head): cached so that steps replaying cached fluxes never read the fronts
  LGAR_S_FD_DZDT,  // dzdt of the wetting_front_free_drainage() front (bmi_lgar.cxx:277-280)
  LGAR_NSCAL
};

// one soil layer of a column class: the reference's soil_properties_ row plus quantities the
// reference recomputes on every call.  Each derived field is produced on the host by exactly the
// expression the reference evaluates (same association order), so reusing it is bit-neutral.
struct LgarLayer {
  double theta_r, theta_e, alpha, n, m, Ksat;
  double de;         // theta_e - theta_r                      (soil_funcs.cxx:149,186)
  double inv_m;      // 1.0/m                                  (soil_funcs.cxx:166)
  double neg_inv_m;  // -1.0/m                                 (soil_funcs.cxx:174)
  double inv_n;      // 1.0/n                                  (soil_funcs.cxx:174)
  double inv_alpha;  // 1.0/alpha                              (soil_funcs.cxx:174)
  double H_c;        // psib*((2+3*lambda)/(1+3*lambda))       (soil_funcs.cxx:123)
  double p_exp;      // 3+1/lambda                             (soil_funcs.cxx:124)
  double theta_min;  // theta(PSI_UPPER_LIM = 1e7)             (lgar.cxx:3222)
  double thick;      // cum[k]-cum[k-1]
  double h_min;      // kept for completeness (unused on the path, soil_funcs.cxx:59-61)
  double ln_alpha;   // log(alpha): (alpha*h)^n is evaluated as exp(n*(ln_alpha + log h)) in the -DLGAR_FAST_POW build
  double slope_max;  // max over h of |d theta / d h| (at (alpha h)^n = m); bounds mass changes in the fast path of the psi searches, never a result
};

// Parameters that the reference keeps per model object and lets a calibrator set through SetValue (src/bmi_lgar.cxx:1032-1049,
// 1403-1420): nonlinear reservoir, preferential-flow split and ponding.  Part of the class row, so every column can carry its own
// set (a calibration run = one column per parameter set).
struct LgarColumnParams {
  double a, b, frac_to_CR, a_slow, b_slow, frac_slow, spf_factor, ponded_depth_max_cm;
};

struct LgarColumnClass {
  int32_t num_layers;
  int32_t invalid_soil_type;         // Q = precip shortcut (bmi_lgar.cxx:145-179)
  double cum[LGAR_MAXL + 1];         // cum_layer_thickness_cm, 1-based (quirk Q1)
  LgarLayer lay[LGAR_MAXL + 1];      // 1-based
  double min_storage;                // sum theta_r*thick     (bmi_lgar.cxx:423-427)
  double max_storage;                // sum theta_e*thick     (lgar.cxx:2431-2436)
  double largest_Ks;                 // max Ksat              (lgar.cxx:2903-2908)
  double psi_50_cm[LGAR_MAXL + 1];   // calc_aet's psi_wp_cm for a head front in layer k (aet.cxx:47-57)
  double initial_theta[LGAR_MAXL + 1];
  double initial_K[LGAR_MAXL + 1];
  LgarColumnParams par;
};

// engine-wide configuration (uniform per engine instance = per "bin")
struct LgarConfig {
  int32_t use_closed_form_G, adaptive_timestep, allow_flux_caching;
  int32_t free_drainage_enabled, free_drainage_to_CR, PET_affects_precip;
  int32_t nint, num_giuh;
  double mbal_tol;
  double timestep_h;            // config "timestep" (= minimum_timestep_h)
  double forcing_resolution_h;
  double initial_psi_cm, wilting_point_psi_cm, field_capacity_psi_cm;
  double giuh[LGAR_MAXG];
};

// device views (all arrays column-fastest; `stride` = padded column count)
struct LgarDeviceState {
  int64_t ncol, stride;
  // fronts [LGAR_W][stride]
  double *depth, *theta, *psi, *K, *dzdt;
  unsigned long long* fflags;   // [LGAR_FW][stride] nibbles layer | to_bottom<<3: front i in word i/16 at bits 4(i%16)..4(i%16)+3
  int32_t* nfront;              // [stride]
  double* scal;                 // [LGAR_NSCAL][stride]
  int32_t* bits;                // [stride] bit0 cache_fluxes, bit1 runoff_in_prev_step, bits 8..15 cache_count
  int32_t* status;              // [stride]
  int32_t* class_id;            // [stride]
  double* giuhq;                // [num_giuh][stride]  (slot N of the reference queue is always 0)
  double* cum;                  // [LGAR_NCUM][stride]
  double* out;                  // [LGAR_NOUT][stride] last step, metres
  const LgarColumnClass* classes;
  const double* ff;             // null, or [LGAR_MAXL+1][stride] frozen_factor per layer (1-based) of the SFT coupling
  // per-step forcing (mm/h)
  const double* precip;
  const double* pet;
  // work list for the active kernel
  int32_t* work;                // [stride]
  int32_t* work_count;          // [2]: count, cursor
  unsigned long long* paths;    // null, or [LGAR_NPATH] path counters (lgarb_set_path_counters)
  int32_t use_ff;               // hour-by-hour kernels: psi searches take the fast path of lgar_search.cuh
};

#endif
