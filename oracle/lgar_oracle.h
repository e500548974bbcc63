/* oracle/lgar_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C11, scalar, one column at a time) of the hot path of NOAA-OWP/LGAR-C:
 * BmiLGAR::Update() and everything under it.  It is the *checker* the
 * CUDA engine is compared against; it is never linked into, imported by, or called from the
 * product library (lgar_c_b200/csrc).  Only tests/, __graft_entry__.smoke() and the cpu_baseline /
 * --impl reference legs of bench.py may use it.
 *
 * Pinning: the oracle is itself checked, step by step, against the UNMODIFIED reference compiled
 * from /root/reference (oracle/_ref/liblgar_ref.so, see oracle/Makefile) and against the
 * reference's own unit-test known answers (tests/main_unit_test_bmi.cxx:456-457,520-522).
 *
 * Deviations from the reference, all deliberate:
 *   - the wetting-front linked list (src/linked_list.cxx) is a fixed array; front_num is the
 *     1-based array position and is never stored;
 *   - abort()/exit()/assert() become bits in lgo_state.status and the column stops updating;
 *   - frozen_factor == 1.0 everywhere (SFT coupling out of scope);
 *   - NGEN build semantics: no "time >= endtime" early break (src/bmi_lgar.cxx:797-803).
 */
#ifndef LGAR_ORACLE_H
#define LGAR_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define LGO_MAX_FRONTS 64
#define LGO_MAX_LAYERS 8
#define LGO_MAX_GIUH 64

/* status bits (a column with any FAIL bit set no longer updates) */
#define LGO_OK 0
#define LGO_FAIL_MBAL 1          /* |local mass balance| > mbal_tol     (bmi_lgar.cxx:785-788 abort) */
#define LGO_FAIL_NEG_RUNOFF 2    /* runoff < 0                          (bmi_lgar.cxx:614-617 abort) */
#define LGO_FAIL_K_NEGATIVE 4    /* K < 0 in dzdt                       (lgar.cxx:2805-2812 abort)   */
#define LGO_FAIL_THETA_ORDER 8   /* theta1 > theta2 in dzdt             (lgar.cxx:2850-2853 exit)    */
#define LGO_FAIL_OVERFLOW 16     /* more fronts than the array holds                                 */
#define LGO_FAIL_LIST 32         /* list walked off its end (the reference would segfault)           */
#define LGO_FAIL_TOP_DEPTH 64    /* head->depth_cm <= 0                 (bmi_lgar.cxx:795 assert)    */
#define LGO_FAIL_TO_BOTTOM 128   /* more to_bottom fronts than layers   (bmi_lgar.cxx:830-834 abort) */
#define LGO_FAIL_OLD_MASS 256    /* old_mass <= 0                       (lgar.cxx:1654 assert)       */
#define LGO_FAIL_MERGE_DEPTH 512 /* merged depth <= 0                   (lgar.cxx:1916 assert)       */
#define LGO_WARN_NEG_FORCING 65536

typedef struct {
  double theta_r, theta_e, alpha, n, m, lambda, psib, h_min, Ksat, theta_wp;
} lgo_soil;

typedef struct {
  double depth, theta, psi, K, dzdt;
  int layer;     /* 1-based */
  int to_bottom; /* 0/1 */
} lgo_front;

/* static description of one column (reference struct lgar_bmi_parameters, all.hxx:102-169) */
typedef struct {
  int num_layers;
  double cum_thickness[LGO_MAX_LAYERS + 1]; /* 1-based, [0] = 0 (quirk Q1) */
  lgo_soil soil[LGO_MAX_LAYERS + 1];        /* soil_properties[layer_soil_type[k]], 1-based */
  double initial_psi_cm;
  double wilting_point_psi_cm, field_capacity_psi_cm;
  int use_closed_form_G, adaptive_timestep, allow_flux_caching;
  int free_drainage_enabled, free_drainage_to_CR, PET_affects_precip;
  int invalid_soil_type;
  double ponded_depth_max_cm;
  double a, b, frac_to_CR, a_slow, b_slow, frac_slow, spf_factor;
  double mbal_tol;
  double timestep_h;           /* config "timestep" in hours; also minimum_timestep_h */
  double forcing_resolution_h;
  int nint;                    /* 120 (lgar.cxx:992) */
  int num_giuh;
  double giuh[LGO_MAX_GIUH];   /* 0-based copy (bmi_lgar.cxx:105-107) */
} lgo_params;

/* path counters: evidence that the tests reach the reference's rare branches (reference file:line in the comments) */
enum {
  LGO_PATH_MERGE = 0,          /* lgar_merge_wetting_fronts, lgar.cxx:1875-1953 (correction type 1) */
  LGO_PATH_CROSS_LAYER,        /* lgar_wetting_fronts_cross_layer_boundary, :1963-2167 (type 2) */
  LGO_PATH_CROSS_BOTTOM,       /* lgar_wetting_front_cross_domain_boundary, :2176-2251 (type 3) */
  LGO_PATH_DRY_OVER_WET,       /* lgar_fix_dry_over_wet_wetting_fronts, :2260-2307 (type 4) */
  LGO_PATH_THETA_R_DELETE,     /* front deleted because theta < theta_r, :1519-1537 */
  LGO_PATH_SATURATED_DEPTH,    /* saturated front deepened until the mass balance closes, :1687-1723 */
  LGO_PATH_CROSS_LAYER_BOTTOM, /* layer crossing that also passes the domain bottom: depth search, :2090-2159 */
  LGO_PATH_SURFICIAL,          /* lgar_create_surficial_front, :2490-2568 */
  LGO_PATH_CLEAN_REDUNDANT,    /* lgar_clean_redundant_fronts deletes a front, :3171-3195 */
  LGO_PATH_SEARCH,             /* lgar_theta_mass_balance iterates, :2990 */
  LGO_PATH_SEARCH_CONVERGED,   /* ... and ends within the tolerance */
  LGO_PATH_SEARCH_STEP_EXIT,   /* ... on |psi - psi_prev| < 1e-15 and factor < 1e-13, :3063 */
  LGO_PATH_SEARCH_NOCHANGE,    /* ... on five trips without mass change, :3070-3076 */
  LGO_PATH_SEARCH_MAXITER,     /* ... on the trip limit, :3078 */
  LGO_PATH_SEARCH_PSI_UPPER,   /* ... on psi > 1e7 after 1000 trips, :3081 */
  LGO_PATH_SEARCH_PSI_ZERO,    /* ... on psi <= 0 twice, :3084 */
  LGO_PATH_CORR_SEARCH,        /* lgar_theta_mass_balance_correction iterates, :3307 */
  LGO_PATH_CACHED_STEP,        /* sub-step that replays cached fluxes, bmi_lgar.cxx:683-695 */
  LGO_PATH_ACTIVE_STEP,        /* sub-step that computes fluxes */
  LGO_NPATH = 24
};

/* everything that survives between Update() calls (SURVEY.md a-S) */
typedef struct {
  int n;
  lgo_front f[LGO_MAX_FRONTS];
  double volon_timestep_cm, precip_previous_timestep_cm;
  double CR_fast_storage_cm, CR_slow_storage_cm;
  double previous_AET, previous_PET, previous_recharge;
  double accumulated_PET, accumulated_free_drainage;
  double timestep_h;
  int cache_fluxes, cache_count, runoff_in_prev_step;
  double giuh_queue[LGO_MAX_GIUH + 1];
  double local_mass_balance;
  /* cumulative volumes for the global balance (lgar.cxx:1162-1185) */
  double volstart_cm, volprecip_cm, volin_cm, volend_cm, volon_cm, volrunoff_cm, volrunoff_giuh_cm;
  double volrech_cm, volAET_cm, volPET_cm, volrunoff_CR_cm, volCRend_cm, volQ_cm, volQ_CR_cm;
  long long timesteps; /* sub-steps taken (lgar_bmi_params.timesteps) */
  int status;
  /* instrumentation (not part of the model) */
  long long n_theta_calls, n_tmb_iters, n_active_substeps;
  /* how often each rare path of the reference ran (same numbering as LGAR_PATH_* in lgar_c_b200/csrc/lgar_types.h) */
  long long n_path[LGO_NPATH];
} lgo_state;

/* order of the 12 per-step scalars, all in metres (bmi_lgar.cxx:882-893):
 * 0 precipitation, 1 potential_evapotranspiration, 2 actual_evapotranspiration, 3 surface_runoff,
 * 4 giuh_runoff, 5 soil_storage, 6 total_discharge, 7 infiltration, 8 percolation,
 * 9 conceptual_reservoir_to_stream_discharge, 10 mass_balance, 11 conceptual_reservoir_storage */
#define LGO_NSCALARS 12

void lgo_soil_derive(lgo_soil* s, double theta_r, double theta_e, double alpha, double n,
                     double Ksat, double wilting_point_psi_cm, int log_mode);
int lgo_read_soil_file(const char* path, int num_soil_types, double wilting_point_psi_cm,
                       int log_mode, lgo_soil* table /* 1-based, num_soil_types+1 entries */);
void lgo_init_state(const lgo_params* p, lgo_state* s);
int lgo_update(const lgo_params* p, lgo_state* s, double precip_mm_per_h, double pet_mm_per_h,
               double out[LGO_NSCALARS]);
int lgo_run_series(const lgo_params* p, lgo_state* s, int nsteps, const double* precip,
                   const double* pet, double* scalars, int* nfronts, int cap, double* fronts,
                   int* flags);
double lgo_global_balance(const lgo_state* s);

/* leaf functions exported for element-wise kernel tests */
double lgo_theta_from_h(double h, const lgo_soil* s);
double lgo_Se_from_h(double h, double alpha, double m, double n);
double lgo_K_from_Se(double Se, double Ks, double m);
double lgo_h_from_Se(double Se, double alpha, double m, double n);
double lgo_Geff(int use_closed_form_G, double theta1, double theta2, const lgo_soil* s, int nint);
double lgo_calc_mass_bal(const double* cum, const lgo_front* f, int n);
double lgo_calc_aet(double PET_cm_per_h, double dt_h, double wp_psi, double fc_psi,
                    const lgo_front* head, const lgo_soil* s);
double lgo_calc_CR_Q(double dt, double a_fast, double a_slow, double b_fast, double b_slow,
                     double frac_slow, double in_cm_per_h, double* fast, double* slow);
double lgo_giuh(double runoff, int n, const double* ord, double* queue);

#ifdef __cplusplus
}
#endif
#endif
