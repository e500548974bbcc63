/* include/lgarb.h -- C ABI of the B200-native batched LGAR/CASAM engine (liblgarb.so).
 *
 * This is the drop-in boundary for the reference's Basic Model Interface, extended with a column
 * batch: one engine advances `n_columns` independent soil columns per Update().  Every entry point
 * names the reference interface it replaces (NOAA-OWP/LGAR-C, file:line).  The reference exports
 * only three C symbols (bmi_model_create / bmi_model_destroy under -DNGEN,
 * include/bmi_lgar.hxx:187-215, and giuh_convolution_integral); everything else is the C++ class
 * `BmiLGAR : bmi::Bmi` (include/bmi_lgar.hxx:30-184), so each BMI virtual gets a C function here
 * taking an opaque handle.  INTEGRATION.md shows the reference-side binding.
 *
 * Conventions
 *   - return value: LGARB_SUCCESS (0 == BMI_SUCCESS) or LGARB_FAILURE (1 == BMI_FAILURE)
 *     (bmi/bmi.hxx:14-15); lgarb_last_error() gives the message.  No exception crosses the ABI and a
 *     numerical failure in one column never aborts the process: see lgarb_get_status().
 *   - units and names of variables are exactly the reference's (src/bmi_lgar.cxx:1186-1210,
 *     1319-1430): inputs in mm/h, per-step outputs in metres.
 *   - batched layout: scalar variables are `double[n_columns]` (int32 for soil_num_wetting_fronts);
 *     the two ragged per-front arrays are returned padded, `double[n_columns][LGARB_MAX_FRONTS]`
 *     with NaN beyond soil_num_wetting_fronts[c]; soil_depth_layers is `double[n_columns][num_layers]`.
 *   - "host" pointers are ordinary (optionally pinned) host memory; "_device" variants take CUDA
 *     device pointers on the engine's device and never stage through the host.
 *   - calls on one handle must be serialised by the caller (same as the reference, which is not
 *     re-entrant per instance); distinct handles are independent.
 *   - There is no CPU fallback: lgarb_initialize fails if no CUDA device is usable.
 */
#ifndef LGARB_H
#define LGARB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LGARB_SUCCESS 0
#define LGARB_FAILURE 1
#define LGARB_MAX_FRONTS 32 /* front capacity per column; overflow sets LGARB_STATUS_FRONT_OVERFLOW */
#define LGARB_MAX_LAYERS 4
#define LGARB_NUM_OUTPUT_SCALARS 12

/* per-column status bits (lgarb_get_status).  The low 16 bits replace the reference's
 * abort()/exit()/assert sites; a column with any of them set is frozen and reports NaN. */
#define LGARB_STATUS_OK 0
#define LGARB_STATUS_MBAL_TOL_EXCEEDED 1   /* src/bmi_lgar.cxx:785-788 abort() */
#define LGARB_STATUS_NEGATIVE_RUNOFF 2     /* src/bmi_lgar.cxx:614-617 abort() */
#define LGARB_STATUS_K_NEGATIVE 4          /* src/lgar.cxx:2805-2812 abort()   */
#define LGARB_STATUS_THETA_ORDER 8         /* src/lgar.cxx:2850-2853 exit()    */
#define LGARB_STATUS_FRONT_OVERFLOW 16     /* > LGARB_MAX_FRONTS wetting fronts */
#define LGARB_STATUS_LIST_CORRUPT 32       /* the reference would dereference NULL */
#define LGARB_STATUS_TOP_DEPTH 64          /* src/bmi_lgar.cxx:795 assert      */
#define LGARB_STATUS_TO_BOTTOM_COUNT 128   /* src/bmi_lgar.cxx:830-834 abort() */
#define LGARB_STATUS_OLD_MASS 256          /* src/lgar.cxx:1654 assert         */
#define LGARB_STATUS_MERGE_DEPTH 512       /* src/lgar.cxx:1916 assert         */
#define LGARB_STATUS_NAN 1024
#define LGARB_STATUS_FAIL_MASK 0xFFFF
#define LGARB_STATUS_WARN_NEGATIVE_FORCING 0x10000 /* src/bmi_lgar.cxx:325-337 warning */

typedef struct lgarb_engine lgarb_engine;

/* ---- life cycle ---------------------------------------------------------------------------- */

/* replaces bmi_model_create(), include/bmi_lgar.hxx:197-200 */
int lgarb_create(lgarb_engine** out);
/* replaces bmi_model_destroy(), include/bmi_lgar.hxx:208-211 */
int lgarb_destroy(lgarb_engine* e);

/* replaces BmiLGAR::Initialize(config_file), src/bmi_lgar.cxx:72-113 -> lgar_initialize,
 * src/lgar.cxx:81-194.  Reads the reference's own key=value[unit] config and soil-parameter file
 * (parser quirks included).  All n_columns columns start identical.  device = CUDA ordinal. */
int lgarb_initialize(lgarb_engine* e, const char* config_file, int64_t n_columns, int device);

/* Batch extension of Initialize: per-column soil types and/or layer thicknesses.
 *   layer_soil_type : int32[n_columns][num_layers] indices into the config's soil file (1-based,
 *                     as `layer_soil_type=` in the config), or NULL to use the config's
 *   layer_thickness_cm : double[n_columns][num_layers], or NULL to use the config's
 * num_layers is the config's layer count.  Columns sharing (types, thicknesses) share one
 * read-only parameter row on the device. */
int lgarb_initialize_columns(lgarb_engine* e, const char* config_file, int64_t n_columns, int device,
                             const int32_t* layer_soil_type, const double* layer_thickness_cm);

/* replaces BmiLGAR::Finalize(), src/bmi_lgar.cxx:1080-1109.  Frees device memory.  The global mass
 * balance the reference prints there (src/lgar.cxx:1162-1216) is available per column through
 * lgarb_get_global_mass_balance() before this call. */
int lgarb_finalize(lgarb_engine* e);

const char* lgarb_last_error(const lgarb_engine* e);

/* ---- time stepping -------------------------------------------------------------------------- */

/* replaces BmiLGAR::Update(), src/bmi_lgar.cxx:133-895: one forcing interval for every column
 * (all internal sub-steps included), using the forcing last set.  Synchronous at the ABI. */
int lgarb_update(lgarb_engine* e);
/* replaces BmiLGAR::UpdateUntil(t), src/bmi_lgar.cxx:898-903 (like the reference: t must be > 0 and
 * is otherwise ignored -- exactly one Update; quirk Q10) */
int lgarb_update_until(lgarb_engine* e, double t);

/* Batch extension: advance n_steps forcing intervals from device-resident series laid out
 * [step][series_stride] (mm/h), with no host round trip between steps.  sinks (nullable) is an
 * array of LGARB_NUM_OUTPUT_SCALARS device pointers, each NULL or a [n_steps][sink_stride] buffer
 * receiving that output variable for every step (order: lgarb_output_scalar_name());
 * d_nfronts_sink (nullable) is an int32 [n_steps][sink_stride] buffer receiving
 * soil_num_wetting_fronts after every step.
 * Asynchronous on the engine's own (non-blocking) stream; lgarb_synchronize() or any get_value waits.
 * The caller must make sure that work it queued on its own streams for these buffers (fills, copies)
 * has completed before the call, and must not touch them before lgarb_synchronize(). */
int lgarb_update_steps_device(lgarb_engine* e, int64_t n_steps, const double* d_precip_mm_per_h,
                              const double* d_pet_mm_per_h, int64_t series_stride,
                              double* const* d_sinks, int64_t sink_stride, int32_t* d_nfronts_sink);
/* The same with HOST buffers (what a framework feeding forcing chunks would call): h_precip / h_pet
 * are [n_steps][n_columns] (mm/h, ideally pinned); output_names lists n_outputs per-step scalar
 * variables and h_outputs[i] receives [n_steps][n_columns] values of output_names[i].  Host<->device
 * copies are part of the call; it returns when the outputs are in place. */
int lgarb_update_steps_host(lgarb_engine* e, int64_t n_steps, const double* h_precip_mm_per_h, const double* h_pet_mm_per_h,
                            int n_outputs, const char* const* output_names, double* const* h_outputs);
/* The same with host arrays wider than this engine's batch: row t of the forcing starts at h_precip + t * forcing_pitch, row t of
 * output i at h_outputs[i] + t * output_pitch (in elements, >= n_columns).  One engine of a multi-device group works on its block
 * of columns of global [step][all columns] arrays this way (include/lgarb_group.h). */
int lgarb_update_steps_host_pitched(lgarb_engine* e, int64_t n_steps, const double* h_precip_mm_per_h, const double* h_pet_mm_per_h,
                                    int64_t forcing_pitch, int n_outputs, const char* const* output_names, double* const* h_outputs,
                                    int64_t output_pitch);
/* Streamed variant (forcing / output streaming; replaces the per-step CSV reader and
 * writer loop of src/bmi_main_lgar.cxx:116-157,183-262 for batched runs).  The series are cut into chunks of
 * chunk_steps hours; the host-to-device copy of chunk k+1 and the device-to-host copy of the outputs of chunk k-1
 * overlap the computation of chunk k (separate copy streams, two staging buffers each way; pinned host memory is
 * needed for the overlap).  forcing_dtype / output_dtype: LGARB_F64 or LGARB_F32 element type of the HOST arrays
 * (gridded forcing products are f32; the engine always computes in f64).  chunk_steps <= 0: one chunk. */
#define LGARB_F64 0
#define LGARB_F32 1
int lgarb_update_steps_host_stream(lgarb_engine* e, int64_t n_steps, int64_t chunk_steps, int forcing_dtype, const void* h_precip_mm_per_h,
                                   const void* h_pet_mm_per_h, int n_outputs, const char* const* output_names, int output_dtype,
                                   void* const* h_outputs);
int lgarb_synchronize(lgarb_engine* e);

/* ---- variable access (BMI getters/setters) -------------------------------------------------- */

/* replaces BmiLGAR::SetValue(name, src), src/bmi_lgar.cxx:1455-1467.  src = host array with one
 * item per column.  Settable: precipitation_rate, potential_evapotranspiration_rate (mm/h), and the calibratable
 * parameters of src/bmi_lgar.cxx:1372-1421 -- per COLUMN: smcmax_k, smcmin_k, van_genuchten_n_k,
 * van_genuchten_alpha_k, hydraulic_conductivity_k (k = 1, 2; the k = 3 names are accepted and ignored exactly
 * like the reference, whose GetVarGrid does not list them, :1133-1141); same value for every column of the batch:
 * field_capacity, ponded_depth_max, a, b, frac_to_CR, a_slow, b_slow, frac_slow, spf_factor.  They take effect at
 * the first Update of a run configured with calib_params=true (update_calibratable_parameters, :916-1078): theta
 * of every front is recomputed from its psi, the change of stored water is reported as volchange_calib by
 * lgarb_get_global_mass_balance, and never afterwards (the reference clears its flag, :187-190).  Millions of
 * columns with millions of parameter sets is the calibration use of the engine.
 * soil_temperature_profile (K): [n_columns][num_cells_temp] (grid 4) when the configuration says sft_coupled=true
 * (cells at the depths of soil_z); the frozen factors of src/lgar.cxx:1109-1149 are recomputed from it before the next
 * step.  Ignored without SFT coupling, like the reference's one unused cell. */
int lgarb_set_value(lgarb_engine* e, const char* name, const void* src);
int lgarb_set_value_range(lgarb_engine* e, const char* name, int64_t col0, int64_t ncol, const void* src);
int lgarb_set_value_device(lgarb_engine* e, const char* name, const void* d_src);
/* replaces BmiLGAR::SetValueAtIndices, src/bmi_lgar.cxx:1470-1490 (inds = column indices) */
int lgarb_set_value_at_indices(lgarb_engine* e, const char* name, const int* inds, int count, const void* src);

/* replaces BmiLGAR::GetValue(name, dest), src/bmi_lgar.cxx:1307-1316.  dest = host buffer of
 * lgarb_get_var_nbytes(name) bytes. */
int lgarb_get_value(lgarb_engine* e, const char* name, void* dest);
int lgarb_get_value_range(lgarb_engine* e, const char* name, int64_t col0, int64_t ncol, void* dest);
int lgarb_get_value_device(lgarb_engine* e, const char* name, void* d_dest);
/* replaces BmiLGAR::GetValuePtr, src/bmi_lgar.cxx:1319-1430: a DEVICE pointer to the engine's own
 * array for a scalar variable (valid until the next update; per-front arrays have no flat pointer) */
int lgarb_get_value_ptr(lgarb_engine* e, const char* name, void** d_ptr);
/* replaces BmiLGAR::GetValueAtIndices, src/bmi_lgar.cxx:1432-1452 (inds = column indices) */
int lgarb_get_value_at_indices(lgarb_engine* e, const char* name, void* dest, const int* inds, int count);

/* ---- metadata (same answers as the reference, scaled by the batch where a size is involved) -- */
int lgarb_get_component_name(const lgarb_engine* e, char* name, int cap);        /* :1493-1497 */
int lgarb_get_input_item_count(const lgarb_engine* e, int* count);               /* :1500-1504 */
int lgarb_get_output_item_count(const lgarb_engine* e, int* count);              /* :1507-1511 */
const char* lgarb_get_input_var_name(const lgarb_engine* e, int i);              /* :1514-1523 */
const char* lgarb_get_output_var_name(const lgarb_engine* e, int i);             /* :1526-1535 */
int lgarb_get_var_grid(const lgarb_engine* e, const char* name, int* grid);      /* :1112-1154 */
int lgarb_get_var_type(const lgarb_engine* e, const char* name, char* type, int cap);     /* :1157-1168 */
int lgarb_get_var_itemsize(const lgarb_engine* e, const char* name, int* size);  /* :1171-1182 */
int lgarb_get_var_units(const lgarb_engine* e, const char* name, char* units, int cap);   /* :1185-1210 */
int lgarb_get_var_nbytes(const lgarb_engine* e, const char* name, int64_t* nbytes);       /* :1213-1222 */
int lgarb_get_var_location(const lgarb_engine* e, const char* name, char* loc, int cap);  /* :1225-1249 */
int lgarb_get_current_time(const lgarb_engine* e, double* t);                    /* :1550-1553 */
int lgarb_get_start_time(const lgarb_engine* e, double* t);                      /* :1538-1541 */
int lgarb_get_end_time(const lgarb_engine* e, double* t);                        /* :1544-1547 */
int lgarb_get_time_step(const lgarb_engine* e, double* dt);                      /* :1562-1565 */
int lgarb_get_time_units(const lgarb_engine* e, char* units, int cap);           /* :1556-1559 */
int lgarb_get_grid_rank(const lgarb_engine* e, int grid, int* rank);             /* :1280-1287 */
int lgarb_get_grid_size(const lgarb_engine* e, int grid, int64_t* size);         /* :1290-1303 */
int lgarb_get_grid_type(const lgarb_engine* e, int grid, char* type, int cap);   /* :1567-1574 */

/* ---- batch-only services ------------------------------------------------------------------- */
int64_t lgarb_get_num_columns(const lgarb_engine* e);
int lgarb_get_num_layers(const lgarb_engine* e);
int lgarb_get_num_giuh_ordinates(const lgarb_engine* e);
const char* lgarb_output_scalar_name(int k); /* k in [0, LGARB_NUM_OUTPUT_SCALARS) */

/* per-column status word (see LGARB_STATUS_*) */
int lgarb_get_status(lgarb_engine* e, int64_t col0, int64_t ncol, int32_t* dest);

/* Raw wetting-front state, what the reference keeps in model_state.head (include/all.hxx:45-56,
 * src/bmi_lgar.cxx:905): arrays [ncol][LGARB_MAX_FRONTS] in cm / cm h^-1, layer 1-based; any
 * destination may be NULL.  nfronts = int32[ncol]. */
int lgarb_get_fronts(lgarb_engine* e, int64_t col0, int64_t ncol, double* depth_cm, double* theta,
                     double* psi_cm, double* K_cm_per_h, double* dzdt_cm_per_h, int32_t* layer_num,
                     int32_t* to_bottom, int32_t* nfronts);

/* the terms of lgar_global_mass_balance(), src/lgar.cxx:1162-1216, per column, cm:
 * out[ncol][16] = volstart, volprecip, volin, volend, volon, volrunoff, volrunoff_giuh, volrech,
 *                 volAET, volPET, volrunoff_CR, volCRend, volQ, volchange_calib, volQ_CR,
 *                 global_error */
int lgarb_get_global_mass_balance(lgarb_engine* e, int64_t col0, int64_t ncol, double* out);
/* the same for every column, computed and left on the device behind the engine's stream: d_out[n_columns][16] */
int lgarb_get_global_mass_balance_device(lgarb_engine* e, double* d_out);

/* The parsed soil table row the engine uses for `soil` (1-based), for checking the parser quirk
 *  against the reference: out[10] = theta_r, theta_e, alpha, n, m, lambda, psib,
 * h_min, Ksat, theta_wp (include/all.hxx:66-79). */
int lgarb_get_soil_row(const lgarb_engine* e, int soil, double* out10);

/* Test hook: route every column through the active kernel (no cached-flux fast path), to check
 * that the two kernels implement the same Update(). */
int lgarb_set_force_all_active(lgarb_engine* e, int on);
/* Scheduling of lgarb_update_steps_device.  3 (default) = wavefront scheduler, see below; 2 = one launch for the whole run, every lane
 * running its own stream of columns and hours as a staged program (converged psi-search iterations);
 * 1 = one launch, warp-owned tiles of columns with per-column time pointers; 0 = one stream+active
 * kernel pair per forcing hour (what lgarb_update does); 3 = wavefront scheduler: the same staged
 * program with one small kernel per stage.  All produce identical results. */
int lgarb_set_window_mode(lgarb_engine* e, int mode);
/* mode 3 (wavefront: one small kernel per stage, contexts parked in HBM, compacted queues) tuning:
 * psi-search trips per SEARCH launch (0 keeps), FRONT/SEARCH pairs per round (0 keeps), and the number of
 * slots in flight below which the remaining columns are finished by one thread each (-1 keeps, 0 = never). */
int lgarb_set_wave_params(lgarb_engine* e, int search_quantum, int front_search_pairs, int finish_below);
/* Debug/profiling aid for the lanes kernel: per column, the number of psi-search trips and the
 * kilo-cycles spent in the other stages during the LAST multi-step launch. */
int lgarb_set_work_trace(lgarb_engine* e, int on);
int lgarb_get_work_trace(lgarb_engine* e, int64_t col0, int64_t ncol, uint32_t* search_trips, uint32_t* other_kcycles);
/* warp-level totals of the last launch for the stages FETCH, HEAD, FRONT, SEARCH, TAIL:
 * out[0..4] cycles, out[5..9] stage turns, out[10..14] lanes that took part */
int lgarb_get_stage_profile(lgarb_engine* e, uint64_t* out15);
/* lifetime (us) and start time (us, mod 2^32) of the first 8192 warps of the last launch: out[0..8191], out[8192..16383] */
int lgarb_get_warp_times(lgarb_engine* e, uint32_t* out16384);
/* number of kernels launched by this engine so far, and columns routed to the active kernel in
 * the last step */
int64_t lgarb_get_launch_count(const lgarb_engine* e);
int64_t lgarb_get_last_active_count(lgarb_engine* e);
/* Per-kernel device timing with CUDA events recorded on the engine's own stream (the only place
 * they can see its kernels).  lgarb_set_profiling(e, 1) starts a fresh accumulation;
 * lgarb_get_profile waits for the stream and returns out[6] = { total ms in the stream kernel, total
 * ms in the active kernel, steps timed, sum over steps of columns routed to the active kernel,
 * sum over steps and columns of the wetting-front count after the step, total ms in window kernels }.  While profiling, each
 * step also copies two counters D2D; leave it off outside measurements. */
int lgarb_set_profiling(lgarb_engine* e, int on);
int lgarb_get_profile(lgarb_engine* e, double* out6);
/* the CUDA stream the engine launches on (cudaStream_t), for callers that time with events */
void* lgarb_get_stream(lgarb_engine* e);
/* FP64 roofline denominator ("not in MEASURED_PEAKS"): dependent-free DFMA chains on every SM of the
 * engine's device for about `ms` milliseconds; *tflops = 2 * DFMAs / time, timed with CUDA events. */
int lgarb_measure_fp64_peak(lgarb_engine* e, double ms, double* tflops);
/* The engine's x^y on the device for n pairs of host doubles (out[i] = x[i]^y[i]): the function that stands in for the libm pow
 * the reference calls (reference src/soil_funcs.cxx:147-187), exposed so that a test can compare it with the host's pow bit for
 * bit.  which = 0: the out-of-line function every cold call site uses, 1: the inlined common case of the psi search with its
 * fall-back to the former. */
int lgarb_debug_pow(lgarb_engine* e, int64_t n, const double* x, const double* y, double* out, int which);
/* The leaf functions of the path on a FROZEN SNAPSHOT of the engine's state, element by element on the device (SURVEY.md section 7
 * step 4): out[i] = fn(a[i], b[i]) for column i < n <= n_columns, with the soil of layer `layer` (1-based, clamped to the column's
 * layer count) of column i's class row:
 *   0 calc_theta_from_h(h = a)             reference src/soil_funcs.cxx:147-150
 *   1 calc_Se_from_h(h = a)                :155-159
 *   2 calc_K_from_Se(Se = a)               :164-167
 *   3 calc_h_from_Se(Se = a)               :172-179
 *   4 calc_Geff(theta1 = a, theta2 = b)    :15-131 (closed form or numeric as the configuration says)
 *   5 lgar_calc_mass_bal of column i's current fronts          src/lgar.cxx:2632-2668
 *   6 calc_aet(PET_cm_per_h = a, dt_h = b) on the current fronts   src/aet.cxx:17-77
 *   7 / 8 / 9 calc_CR_Q(dt = a, input_cm_per_h = b) from the column's current reservoir storages: discharge / fast storage after /
 *             slow storage after (the engine's state is not modified)     src/conceptual_reservoir.cxx:7-45
 *   10 / 11   the fast / slow storage that call starts from, as it stands in the state (cm)
 * The GIUH convolution is fused into the step epilogue and is checked through whole steps (giuh_runoff of every golden case). */
int lgarb_debug_functions(lgarb_engine* e, int fn, int layer, int64_t n, const double* a, const double* b, double* out);
/* Path counters: how often the rare branches of the reference ran since they were switched on (merge, layer crossing, domain
 * bottom crossing, dry-over-wet, theta < theta_r deletion, saturated-front depth loop, ..., every exit of the psi search; the
 * LGAR_PATH_* numbering of lgar_c_b200/csrc/lgar_types.h = LGO_PATH_* of oracle/lgar_oracle.h; reference src/lgar.cxx:1519-1537,
 * 1687-1723,1875-2307,3063-3086,3116-3195).  Evidence for tests, off by default (one predicated atomic per rare event when on). */
/* Stage timing of the multi-step verbs: with it on, every launch of a stage kernel (FETCH, HEAD, FRONT, SEARCH, TAIL, CORR) and of
 * the finishing kernel is bracketed by a CUDA event pair on the stream it is launched on; lgarb_get_stage_times returns the
 * accumulated milliseconds and launch counts per stage (7 entries each, the finishing kernel last) and clears nothing.  The
 * correction searches run on a side stream beside the main one, so the sum of the stage times may exceed the window's time. */
int lgarb_set_stage_timing(lgarb_engine* e, int on);
int lgarb_get_stage_times(lgarb_engine* e, double* ms7, int64_t* launches7);
int lgarb_set_path_counters(lgarb_engine* e, int on);
int lgarb_get_path_counters(lgarb_engine* e, uint64_t* out, int n);

#ifdef __cplusplus
}
#endif
#endif /* LGARB_H */
