// lgar_c_b200/csrc/lgar_engine.cu -- host side of liblgarb.so: the reference's BMI verbs over a
// column batch (C ABI declared in include/lgarb.h).
//
//   * config reader      mirrors InitFromConfigFile        (reference src/lgar.cxx:231-1016)
//   * soil-table reader  mirrors lgar_read_vG_param_file   (reference src/lgar.cxx:2674-2758, quirk Q2)
//   * initial fronts     mirrors InitializeWettingFronts   (reference src/lgar.cxx:1024-1064)
//   * Update/GetValue/SetValue/metadata mirror BmiLGAR      (reference src/bmi_lgar.cxx)
//
// The engine owns all device memory (fixed-capacity SoA, lgar_types.h) and a CUDA stream.  There is
// no CPU path: without a usable CUDA device initialisation fails.
#include <cuda_runtime.h>
#include <cuda_profiler_api.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/lgarb.h"
#include "lgar_kernels.h"
#include "lgar_types.h"

#include "lgar_config.hpp"

using namespace lgarhost;

namespace {

#define CUDA_TRY(expr)                                                                                     \
  do {                                                                                                     \
    cudaError_t _err = (expr);                                                                             \
    if (_err != cudaSuccess) throw std::runtime_error(std::string(#expr) + ": " + cudaGetErrorString(_err)); \
  } while (0)

const char* kOutNames[LGAR_NOUT] = {
    "precipitation", "potential_evapotranspiration", "actual_evapotranspiration", "surface_runoff", "giuh_runoff",
    "soil_storage", "total_discharge", "infiltration", "percolation", "conceptual_reservoir_to_stream_discharge",
    "mass_balance", "conceptual_reservoir_storage"};
// bmi_lgar.hxx:34-54
const char* kInputNames[3] = {"precipitation_rate", "potential_evapotranspiration_rate", "soil_temperature_profile"};
const char* kOutputNames[15] = {"soil_moisture_wetting_fronts", "soil_depth_layers", "soil_depth_wetting_fronts", "soil_num_wetting_fronts",
                                "precipitation", "potential_evapotranspiration", "actual_evapotranspiration", "surface_runoff",
                                "giuh_runoff", "soil_storage", "total_discharge", "infiltration", "percolation",
                                "conceptual_reservoir_to_stream_discharge", "mass_balance"};

int out_index(const std::string& name) {
  for (int k = 0; k < LGAR_NOUT; k++)
    if (name == kOutNames[k]) return k;
  return -1;
}

}  // namespace

struct StreamBufs;
void free_stream_bufs(StreamBufs*);

struct lgarb_engine {
  bool initialized = false;
  StreamBufs* sbufs = nullptr;  // staging of lgarb_update_steps_host_stream, kept between calls
  // calibratable parameters (src/bmi_lgar.cxx:916-1078): set through SetValue before the first Update
  std::vector<int32_t> class_of;          // [ncol] class row of every column
  std::vector<int32_t> col_types;         // [ncol][L] soil types (empty: the configuration's for every column)
  std::vector<double> col_thick;          // [ncol][L] layer thicknesses (empty: the configuration's)
  bool calib_pending = false, calib_touched = false;
  std::vector<double> calib_soil[15];     // [layer 1..3][smcmax, smcmin, vg_n, vg_alpha, Ksat] -> [ncol], allocated when first set
  double calib_glob[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // field_capacity, ponded_depth_max, a, b, frac_to_CR, a_slow, b_slow, frac_slow, spf_factor: what Initialize() gives every column
  std::vector<double> calib_glob_col[9];  // the same per column ([ncol], allocated when first set): every column may carry its own set
  std::vector<double> volchange_calib;    // [ncol] cm, change of stored water by the parameter update
  // SFT coupling (sft_coupled=true): soil temperature profile per column -> frozen_factor per layer (lgar.cxx:1109-1149)
  std::vector<double> soil_temperature;   // [ncol][num_cells_temp] K, zero until set (lgar.cxx:947)
  double* d_ff = nullptr;                 // [LGAR_MAXL+1][stride]
  bool ff_dirty = false;
  int device = 0;
  bool has_device = false;  // `device` is the CUDA device this engine's memory and streams live on
  int64_t ncol = 0, stride = 0;
  int num_layers = 0;
  ParsedConfig pc;
  LgarConfig cfg;
  std::vector<SoilRow> soils;  // 1-based
  std::vector<LgarColumnClass> classes;
  LgarDeviceState d;
  std::vector<void*> allocs;
  double *d_precip = nullptr, *d_pet = nullptr, *d_stage = nullptr;
  LgarColumnClass* d_classes = nullptr;
  double* volstart = nullptr;  // [stride] initial storage (cm)
  int32_t* d_stage_i = nullptr;
  void* h_pinned = nullptr;
  size_t h_pinned_bytes = 0;
  cudaStream_t stream = nullptr;
  int active_blocks = 0;
  int window_blocks = 0;
  int use_window = 3;          // multi-step runs: 3 = six stage kernels with queues between them, 1 = warp-tile window kernel, 0 = hour-by-hour kernel pairs
  unsigned int* d_work_trace = nullptr;  // [2][stride], allocated by lgarb_set_work_trace
  double* d_host_series = nullptr;  // staging of lgarb_update_steps_host
  size_t host_series_cap = 0;
  // wavefront scheduler (mode 3)
  LgarWave wave;
  bool wave_ready = false;
  int wave_quantum = 64, wave_pairs = 4, wave_corr_pairs = 1;
  // correction searches on a side stream: the CORR kernel of a round (few slots, a long serial FP64 chain per slot) runs
  // beside FETCH/HEAD/FRONT/SEARCH of the same round and is joined before TAIL; 0 = CORR and a second TAIL at the end of
  // the round on the main stream
  int wave_corr_async = 1;
  cudaStream_t side_stream = nullptr;
  cudaEvent_t side_ev[2] = {nullptr, nullptr};
  // Forcing on its way (lgarb_update_steps_host): the series is copied in chunks on a stream of its own while the window runs;
  // chunk k is complete when ev[k] is, and the window may then use the hours below upto[k].  Null: everything is present.
  struct Frontier {
    std::vector<cudaEvent_t> ev;
    std::vector<int> upto;
    size_t next = 0;
  };
  Frontier* frontier = nullptr;
  cudaStream_t h2d_stream = nullptr;
  int host_chunk_hours = 24;   // lgarb_update_steps_host: hours per chunk of the forcing copy (0: copy everything first)
  int64_t frontier_waits = 0;  // times the scheduler had to wait for forcing
  int wave_finish_below = 16384;
  // Rounds with at most this many slots in flight run their psi searches through the fast path (lgar_search.cuh).  0: never --
  // measured (profiles/r2_call5_call6_summary.md): a round lasts as long as its longest search, and the fast path runs every search
  // to its end inside one launch where the plain walk parks a 10^4-trip search after `quantum` trips: 4.09e8 with 65 536 against
  // 4.74e8 with 0.
  int wave_ff_below = 0;
  // The finishing kernel keeps its lane records in shared memory (lgar_finish_smem.cu); 0: round 1's kernel, one thread per slot
  // out of thread-private records.  Measured: 4.79e8 -> 5.00e8 (call of 480 h), 5.29e8 -> 5.48e8 (1 460 h).
  int finish_smem = 1;
  int finish_spw = 16;           // records per warp of that kernel: 32 ... 2
  int finish_quantum = 64;       // search trips per turn of its warps
  int finish_ff = 1;             // finishing kernel: searches through the fast path (no difference for long windows; the hour-by-hour Update ends on the slowest search of the batch)
  unsigned long long* d_paths = nullptr;  // [LGAR_NPATH] path counters, allocated by lgarb_set_path_counters
  int finish_lane_stride = 4;  // finishing kernel: one slot per this many lanes
  int sm_count = 0;            // multiprocessors of the device: grids are sized in multiples of it
  // lgarb_update: batches of at least this many columns hand their active columns to the stage kernels (measured on 1.25 M columns,
  // profiles/r2_call10_15_summary.md: 23.0 ms per hour with the stream/active kernel pair, 9.8 ms this way)
  int update_wave_min_cols = 100000;
  int update_wave_slots = 262144;      // ... through this many slots
  int update_wave_finish_below = 40000;  // ... with the finishing kernel taking over below this many slots in flight
  int wave_windows = 0;  // windows run so far (profiling range selection)
  int wave_escalate_below = 8192;
  int wave_corr_quantum = 64;  // correction-search trips per launch (scaled with the psi-search quantum when that escalates)
  int64_t wave_max_slots = 16777216;  // 2.8 KB of context per slot
  int64_t wave_rounds = 0;
  int32_t* d_tile_counter = nullptr;
  cudaEvent_t win_ev[2] = {nullptr, nullptr};
  double prof_ms_window = 0;
  int64_t prof_windows = 0;
  int force_all_active = 0;
  int64_t steps = 0;
  int64_t launches = 0;
  // stage timing (lgarb_set_stage_timing): an event pair round every stage launch of the window scheduler, resolved lazily
  bool stage_timing = false;
  struct StageEv { cudaEvent_t a, b; int stage; };
  std::vector<StageEv> stage_evs;
  double stage_ms[LGAR_NSTAGE + 1] = {0, 0, 0, 0, 0, 0, 0};   // FETCH, HEAD, FRONT, SEARCH, TAIL, CORR, finishing kernel
  int64_t stage_launches[LGAR_NSTAGE + 1] = {0, 0, 0, 0, 0, 0, 0};
  // profiling (lgarb_set_profiling): event triples per step, resolved lazily
  bool profiling = false;
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  double prof_ms_stream = 0, prof_ms_active = 0;
  int64_t prof_steps = 0;
  unsigned long long* d_prof_acc = nullptr;  // [2]: sum nfront, sum active count
  double time_s = 0.0;
  std::string last_error;

  template <typename T>
  T* dalloc(size_t n, bool zero = true) {
    void* p = nullptr;
    CUDA_TRY(cudaMalloc(&p, n * sizeof(T)));
    allocs.push_back(p);
    if (zero) CUDA_TRY(cudaMemsetAsync(p, 0, n * sizeof(T), stream));
    return (T*)p;
  }
  void* pinned(size_t bytes) {
    if (bytes > h_pinned_bytes) {
      if (h_pinned) cudaFreeHost(h_pinned);
      h_pinned = nullptr;
      CUDA_TRY(cudaMallocHost(&h_pinned, bytes));
      h_pinned_bytes = bytes;
    }
    return h_pinned;
  }
  void release() {
    if (stream) cudaStreamSynchronize(stream);
    d_work_trace = nullptr;
    d_paths = nullptr;
    wave_ready = false;
    if (d_host_series) cudaFree(d_host_series);
    d_host_series = nullptr;
    host_series_cap = 0;
    free_stream_bufs(sbufs);
    sbufs = nullptr;
    for (auto& se : stage_evs) {
      cudaEventDestroy(se.a);
      cudaEventDestroy(se.b);
    }
    stage_evs.clear();
    stage_timing = false;
    for (cudaEvent_t x : ev_pool) cudaEventDestroy(x);
    for (int i = 0; i < 2; i++) {
      if (win_ev[i]) cudaEventDestroy(win_ev[i]);
      win_ev[i] = nullptr;
    }
    ev_pool.clear();
    ev_used = 0;
    profiling = false;
    for (void* p : allocs) cudaFree(p);
    allocs.clear();
    if (h_pinned) cudaFreeHost(h_pinned);
    h_pinned = nullptr;
    h_pinned_bytes = 0;
    for (int i = 0; i < 2; i++) {
      if (side_ev[i]) cudaEventDestroy(side_ev[i]);
      side_ev[i] = nullptr;
    }
    if (side_stream) cudaStreamDestroy(side_stream);
    side_stream = nullptr;
    if (h2d_stream) cudaStreamDestroy(h2d_stream);
    h2d_stream = nullptr;
    if (stream) cudaStreamDestroy(stream);
    stream = nullptr;
    initialized = false;
    has_device = false;
  }
};

namespace {

void do_initialize_body(lgarb_engine* e, const char* config_file, int64_t ncol, int device, const int32_t* types_in, const double* thick_in);
// A failed initialisation (out of memory at a large n_columns is the likely one) must not keep the stream and the allocations it
// already made: the engine is released, so that a retry with a smaller batch starts from a clean device.
void do_initialize(lgarb_engine* e, const char* config_file, int64_t ncol, int device, const int32_t* types_in, const double* thick_in) {
  try {
    do_initialize_body(e, config_file, ncol, device, types_in, thick_in);
  } catch (...) {
    e->release();
    throw;
  }
}
void do_initialize_body(lgarb_engine* e, const char* config_file, int64_t ncol, int device, const int32_t* types_in, const double* thick_in) {
  e->release();  // tolerates an engine that was never (or only partly) initialised
  if (ncol <= 0) throw std::runtime_error("n_columns must be positive");
  if (ncol > 0x7FFFFFF0LL) throw std::runtime_error("n_columns must fit in int32");
  int ndev = 0;
  cudaError_t derr = cudaGetDeviceCount(&ndev);
  if (derr != cudaSuccess || ndev == 0)
    throw std::runtime_error(std::string("no usable CUDA device (this engine has no CPU path): ") + cudaGetErrorString(derr));
  if (device < 0 || device >= ndev) throw std::runtime_error("invalid CUDA device ordinal");
  CUDA_TRY(cudaSetDevice(device));
  e->device = device;
  e->has_device = true;
  e->pc = parse_config(config_file);
  const ParsedConfig& pc = e->pc;
  const int L = (int)pc.layer_thickness.size();
  if (L < 1 || L > LGAR_MAXL) throw std::runtime_error("number of soil layers must be 1.." + std::to_string(LGAR_MAXL));
  if ((int)pc.layer_soil_type.size() != L) throw std::runtime_error("layer_soil_type and layer_thickness differ in length");
  e->num_layers = L;
  int nread = read_soil_file(pc.soil_params_file, pc.num_soil_types, pc.wilting_point_psi, pc.log_mode, e->soils, config_file);

  e->cfg = make_engine_config(pc);
  LgarConfig& g = e->cfg;

  // column classes: one parameter row per distinct (soil types, thicknesses)
  std::vector<int32_t> class_of((size_t)ncol);
  e->classes.clear();
  std::map<std::vector<double>, int> seen;
  std::vector<int> types(L);
  std::vector<double> thick(L);
  for (int64_t c = 0; c < ncol; c++) {
    std::vector<double> key(2 * L);
    for (int k = 0; k < L; k++) {
      types[k] = types_in ? types_in[c * L + k] : pc.layer_soil_type[k];
      thick[k] = thick_in ? thick_in[c * L + k] : pc.layer_thickness[k];
      key[k] = types[k];
      key[L + k] = thick[k];
    }
    auto it = seen.find(key);
    if (it == seen.end()) {
      check_column_class(pc, nread, L, types.data(), thick.data());
      int id = (int)e->classes.size();
      e->classes.push_back(make_class(pc, e->soils, L, types.data(), thick.data()));
      it = seen.emplace(key, id).first;
    }
    class_of[c] = it->second;
    if (!types_in && !thick_in) {  // homogeneous batch: every column is class 0
      for (int64_t c2 = 1; c2 < ncol; c2++) class_of[c2] = 0;
      break;
    }
  }

  e->class_of = class_of;
  e->col_types.clear();
  e->col_thick.clear();
  if (types_in) e->col_types.assign(types_in, types_in + (size_t)ncol * L);
  if (thick_in) e->col_thick.assign(thick_in, thick_in + (size_t)ncol * L);
  e->soil_temperature.clear();
  e->d_ff = nullptr;
  e->ff_dirty = false;
  e->calib_pending = pc.calib_params;  // lgar.cxx:742: applied once, at the first Update
  e->calib_touched = false;
  for (auto& v : e->calib_soil) v.clear();
  for (auto& v : e->calib_glob_col) v.clear();
  e->volchange_calib.clear();
  {  // lgar.cxx:105-113
    const double g0[9] = {pc.field_capacity_psi, pc.ponded_depth_max, pc.a, pc.b, pc.frac_to_CR, pc.a_slow, pc.b_slow, pc.frac_slow, pc.spf_factor};
    for (int i = 0; i < 9; i++) e->calib_glob[i] = g0[i];
  }

  // device memory
  CUDA_TRY(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  e->ncol = ncol;
  e->stride = (ncol + 31) / 32 * 32;
  const size_t st = (size_t)e->stride;
  LgarDeviceState& d = e->d;
  memset(&d, 0, sizeof(d));
  d.ncol = ncol;
  d.stride = e->stride;
  d.depth = e->dalloc<double>(st * LGAR_W);
  d.theta = e->dalloc<double>(st * LGAR_W);
  d.psi = e->dalloc<double>(st * LGAR_W);
  d.K = e->dalloc<double>(st * LGAR_W);
  d.dzdt = e->dalloc<double>(st * LGAR_W);
  d.fflags = e->dalloc<unsigned long long>(st * LGAR_FW);
  d.nfront = e->dalloc<int32_t>(st);
  d.scal = e->dalloc<double>(st * LGAR_NSCAL);
  d.bits = e->dalloc<int32_t>(st);
  d.status = e->dalloc<int32_t>(st);
  d.class_id = e->dalloc<int32_t>(st);
  d.giuhq = e->dalloc<double>(st * (size_t)std::max(1, g.num_giuh - 1));
  d.cum = e->dalloc<double>(st * LGAR_NCUM);
  d.out = e->dalloc<double>(st * LGAR_NOUT);
  d.work = e->dalloc<int32_t>(st);
  d.work_count = e->dalloc<int32_t>(2);
  e->d_precip = e->dalloc<double>(st);
  e->d_pet = e->dalloc<double>(st);
  e->volstart = e->dalloc<double>(st);
  e->d_prof_acc = e->dalloc<unsigned long long>(2);
  e->d_tile_counter = e->dalloc<int32_t>(1);
  e->d_stage = e->dalloc<double>(st * LGAR_W);
  e->d_stage_i = e->dalloc<int32_t>(st);
  e->d_classes = e->dalloc<LgarColumnClass>(e->classes.size(), false);
  d.classes = e->d_classes;
  d.ff = nullptr;
  if (pc.sft_coupled) {
    e->soil_temperature.assign((size_t)ncol * pc.soil_z.size(), 0.0);
    e->d_ff = e->dalloc<double>(st * (LGAR_MAXL + 1));
    d.ff = e->d_ff;
    e->ff_dirty = true;  // the first Update computes the factors from the (zero) profile like the reference does
  }
  d.precip = e->d_precip;
  d.pet = e->d_pet;
  CUDA_TRY(cudaMemcpyAsync(e->d_classes, e->classes.data(), e->classes.size() * sizeof(LgarColumnClass), cudaMemcpyHostToDevice, e->stream));
  CUDA_TRY(cudaMemcpyAsync(d.class_id, class_of.data(), (size_t)ncol * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));

  // initial state: InitializeWettingFronts (lgar.cxx:1024-1064) + lgar_initialize (lgar.cxx:172-193).
  // Forcing defaults to -1 so that a missing SetValue is visible (lgar.cxx:172-175).
  {
    std::vector<double> h((size_t)ncol);
    std::vector<double> buf((size_t)ncol * 5 * L);
    auto plane = [&](int which, int k) { return buf.data() + ((size_t)which * L + k) * (size_t)ncol; };
    std::vector<unsigned long long> fl((size_t)ncol);
    std::vector<int32_t> nf((size_t)ncol);
    std::vector<double> stor((size_t)ncol), tsh((size_t)ncol, pc.timestep_h);
    for (int64_t c = 0; c < ncol; c++) {
      const LgarColumnClass& cc = e->classes[class_of[c]];
      unsigned long long f = 0;
      if (!cc.invalid_soil_type) {
        for (int k = 0; k < L; k++) {
          plane(0, k)[c] = cc.cum[k + 1];
          plane(1, k)[c] = cc.initial_theta[k + 1];
          plane(2, k)[c] = pc.initial_psi;
          plane(3, k)[c] = cc.initial_K[k + 1];
          plane(4, k)[c] = 0.0;
          f |= (unsigned long long)(((k + 1) & 7) | 8) << (4 * k);
        }
        nf[c] = L;
        stor[c] = class_initial_storage(cc);
      } else {
        nf[c] = 0;
        stor[c] = 0.0;
      }
      fl[c] = f;
    }
    double* dst[5] = {d.depth, d.theta, d.psi, d.K, d.dzdt};
    for (int w = 0; w < 5; w++)
      for (int k = 0; k < L; k++)
        CUDA_TRY(cudaMemcpyAsync(dst[w] + (size_t)k * st, plane(w, k), (size_t)ncol * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(cudaMemcpyAsync(d.fflags, fl.data(), (size_t)ncol * sizeof(unsigned long long), cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(cudaMemcpyAsync(d.nfront, nf.data(), (size_t)ncol * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(cudaMemcpyAsync(d.scal + (size_t)LGAR_S_STORAGE * st, stor.data(), (size_t)ncol * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(cudaMemcpyAsync(e->volstart, stor.data(), (size_t)ncol * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(cudaMemcpyAsync(d.scal + (size_t)LGAR_S_TIMESTEP_H * st, tsh.data(), (size_t)ncol * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    std::vector<int32_t> bits((size_t)ncol, 1 << 8);  // cache_count = 1 (all.hxx:165)
    CUDA_TRY(cudaMemcpyAsync(d.bits, bits.data(), (size_t)ncol * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
    std::vector<double> minus1((size_t)ncol, -1.0);
    CUDA_TRY(cudaMemcpyAsync(e->d_precip, minus1.data(), (size_t)ncol * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(cudaMemcpyAsync(e->d_pet, minus1.data(), (size_t)ncol * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
  }

  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  int per_sm = lgar_active_kernel_max_blocks_per_sm();
  if (per_sm < 1) per_sm = 1;
  e->sm_count = prop.multiProcessorCount;
  e->active_blocks = prop.multiProcessorCount * per_sm;  // persistent: a whole number of waves of the SM count
  int wper_sm = lgar_window_kernel_max_blocks_per_sm();
  if (wper_sm < 1) wper_sm = 1;
  e->window_blocks = prop.multiProcessorCount * wper_sm;
  e->steps = 0;
  e->launches = 0;
  e->time_s = 0.0;
  e->initialized = true;
}

// ---- calibratable parameters ---------------------------------------------------
// index into calib_soil for "smcmax_2" etc. (src/bmi_lgar.cxx:1372-1402), -1 if `name` is not one of them
int calib_soil_index(const std::string& name) {
  static const char* stem[5] = {"smcmax_", "smcmin_", "van_genuchten_n_", "van_genuchten_alpha_", "hydraulic_conductivity_"};
  for (int f = 0; f < 5; f++) {
    const size_t n = strlen(stem[f]);
    if (name.size() == n + 1 && name.compare(0, n, stem[f]) == 0 && name[n] >= '1' && name[n] <= '3') return (name[n] - '1') * 5 + f;
  }
  return -1;
}
int calib_glob_index(const std::string& name) {  // src/bmi_lgar.cxx:1404-1421
  static const char* nm[9] = {"field_capacity", "ponded_depth_max", "a", "b", "frac_to_CR", "a_slow", "b_slow", "frac_slow", "spf_factor"};
  for (int i = 0; i < 9; i++)
    if (name == nm[i]) return i;
  return -1;
}
// value Initialize() gives a per-layer calibratable parameter of column c (lgar.cxx:131-153): the soil row of the layer
double calib_soil_initial(const lgarb_engine* e, int idx, int64_t c) {
  const int layer = idx / 5 + 1, f = idx % 5;
  const LgarColumnClass& cc = e->classes[e->class_of[c]];
  if (cc.invalid_soil_type || layer > cc.num_layers) return 0.0;
  const LgarLayer& l = cc.lay[layer];
  return f == 0 ? l.theta_e : f == 1 ? l.theta_r : f == 2 ? l.n : f == 3 ? l.alpha : l.Ksat;
}

// update_calibratable_parameters (src/bmi_lgar.cxx:916-1078), once, before the first forcing step of a run whose
// configuration says calib_params=true: every column gets the parameter set that was written through SetValue,
// theta of every front is recomputed from its psi, and the change of stored water is booked in volchange_calib.
// Host side (libm, the reference's own expressions): the result is what the reference computes, bit for bit.
void apply_calibration(lgarb_engine* e) {
  if (!e->calib_pending) return;
  e->calib_pending = false;
  const ParsedConfig pc0 = e->pc;
  if (!e->calib_touched && !pc0.log_mode) return;  // same parameters: theta(psi) reproduces the initial theta, nothing changes
  ParsedConfig& pc = e->pc;
  const int L = e->num_layers;
  const int64_t ncol = e->ncol;
  const size_t st = (size_t)e->stride;
  // parameters that do not depend on the layer (:1031-1049): per column, written into a scratch copy of the configuration from
  // which the column's class row is made
  ParsedConfig pcc = pc0;
  auto glob = [&](int gi, int64_t c) { return e->calib_glob_col[gi].empty() ? e->calib_glob[gi] : e->calib_glob_col[gi][(size_t)c]; };
  auto column_config = [&](int64_t c) {
    pcc.field_capacity_psi = glob(0, c);
    pcc.ponded_depth_max = glob(1, c);
    pcc.a = glob(2, c);
    pcc.b = glob(3, c);
    pcc.frac_to_CR = glob(4, c);
    pcc.spf_factor = glob(8, c);
    if (pc0.frac_slow) {
      pcc.a_slow = glob(5, c);
      pcc.b_slow = glob(6, c);
      pcc.frac_slow = glob(7, c);
    }
    if (pc0.log_mode) {
      pcc.a = pow(10.0, glob(2, c));
      if (pc0.frac_slow) pcc.a_slow = pow(10.0, glob(5, c));
    }
  };
  column_config(0);
  // the batch-level copy keeps column 0's values (what the metadata verbs report when every column has the same set)
  pc.field_capacity_psi = pcc.field_capacity_psi; pc.ponded_depth_max = pcc.ponded_depth_max; pc.a = pcc.a; pc.b = pcc.b;
  pc.frac_to_CR = pcc.frac_to_CR; pc.spf_factor = pcc.spf_factor; pc.a_slow = pcc.a_slow; pc.b_slow = pcc.b_slow; pc.frac_slow = pcc.frac_slow;
  e->cfg.field_capacity_psi_cm = pc.field_capacity_psi;

  CUDA_TRY(cudaStreamSynchronize(e->stream));
  std::vector<int32_t> nf((size_t)ncol);
  std::vector<unsigned long long> fl((size_t)ncol * LGAR_FW);  // [word][column]
  CUDA_TRY(cudaMemcpy(nf.data(), e->d.nfront, (size_t)ncol * sizeof(int32_t), cudaMemcpyDeviceToHost));
  for (int wd = 0; wd < LGAR_FW; wd++)
    CUDA_TRY(cudaMemcpy(fl.data() + (size_t)wd * ncol, e->d.fflags + (size_t)wd * st, (size_t)ncol * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  auto nibble = [&](int64_t c, int i) { return (int)((fl[(size_t)(i >> 4) * ncol + c] >> (4 * (i & 15))) & 0xF); };
  int maxnf = 0;
  for (int64_t c = 0; c < ncol; c++) maxnf = std::max(maxnf, (int)nf[c]);
  std::vector<double> depth((size_t)maxnf * ncol), theta((size_t)maxnf * ncol), psi((size_t)maxnf * ncol);
  for (int i = 0; i < maxnf; i++) {
    CUDA_TRY(cudaMemcpy(depth.data() + (size_t)i * ncol, e->d.depth + (size_t)i * st, (size_t)ncol * sizeof(double), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(theta.data() + (size_t)i * ncol, e->d.theta + (size_t)i * st, (size_t)ncol * sizeof(double), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(psi.data() + (size_t)i * ncol, e->d.psi + (size_t)i * st, (size_t)ncol * sizeof(double), cudaMemcpyDeviceToHost));
  }
  auto mass = [&](int64_t c, const double* cum) {  // lgar_calc_mass_bal, lgar.cxx:2632-2668
    double sum = 0.0;
    const int n = nf[c];
    for (int i = 0; i < n; i++) {
      const int li = nibble(c, i) & 7;
      const double base = cum[li - 1];
      const double th = theta[(size_t)i * ncol + c], dp = depth[(size_t)i * ncol + c];
      if (i + 1 < n) {
        const int ln = nibble(c, i + 1) & 7;
        if (ln == li) sum += (dp - base) * (th - theta[(size_t)(i + 1) * ncol + c]);
        else sum += (dp - base) * th;
      } else {
        sum += th * (dp - base);
      }
    }
    return sum;
  };
  std::vector<LgarColumnClass> classes;
  std::map<std::vector<double>, int> seen;
  std::vector<int32_t> class_of((size_t)ncol);
  std::vector<double> storage((size_t)ncol, 0.0);
  e->volchange_calib.assign((size_t)ncol, 0.0);
  std::vector<int> types(L);
  std::vector<double> thick(L);
  std::vector<SoilRow> rows((size_t)L + 1);
  for (int64_t c = 0; c < ncol; c++) {
    const LgarColumnClass& old = e->classes[e->class_of[c]];
    for (int k = 0; k < L; k++) {
      types[k] = e->col_types.empty() ? pc.layer_soil_type[k] : e->col_types[(size_t)c * L + k];
      thick[k] = e->col_thick.empty() ? pc.layer_thickness[k] : e->col_thick[(size_t)c * L + k];
    }
    std::vector<double> key;
    column_config(c);
    if (old.invalid_soil_type) {  // no fronts, precipitation passes through: the class stays what it is
      key.assign(1, -1.0 - e->class_of[c]);
    } else {
      // the reference walks the fronts top down and overwrites soil_properties[soil type of the front's layer] with
      // the parameters of that LAYER (:929-990): layers that share a soil type end up with the parameters of the
      // deepest of them, but a front's theta is computed with what the row holds when the walk reaches it
      for (int k = 1; k <= L; k++) rows[k] = e->soils[types[k - 1]];
      const double before = mass(c, old.cum);
      for (int i = 0; i < nf[c]; i++) {
        const int layer = nibble(c, i) & 7;
        if (layer <= 3) {
          auto val = [&](int f) {
            const int idx = (layer - 1) * 5 + f;
            return e->calib_soil[idx].empty() ? calib_soil_initial(e, idx, c) : e->calib_soil[idx][(size_t)c];
          };
          SoilRow r = rows[layer];
          r.theta_e = val(0);
          r.theta_r = val(1);
          r.n = val(2);
          r.m = 1.0 - 1.0 / r.n;
          r.alpha = val(3);
          r.Ksat = val(4);
          if (pc.log_mode) {
            r.alpha = pow(10.0, val(3));
            r.Ksat = pow(10.0, val(4));
          }
          for (int k = 1; k <= L; k++)
            if (types[k - 1] == types[layer - 1]) rows[k] = r;  // one row per soil TYPE in the reference
        }
        theta[(size_t)i * ncol + c] = theta_from_h(psi[(size_t)i * ncol + c], rows[layer]);
      }
      const double after = mass(c, old.cum);
      e->volchange_calib[(size_t)c] = after - before;
      storage[(size_t)c] = after;
      for (int k = 0; k < L; k++) key.push_back(thick[k]);
      {
        const double gk[9] = {pcc.field_capacity_psi, pcc.ponded_depth_max, pcc.a, pcc.b, pcc.frac_to_CR, pcc.a_slow, pcc.b_slow, pcc.frac_slow, pcc.spf_factor};
        key.insert(key.end(), gk, gk + 9);
      }
      for (int k = 1; k <= L; k++) {
        const SoilRow& r = rows[k];
        const double f[10] = {r.theta_r, r.theta_e, r.alpha, r.n, r.m, r.lambda, r.psib, r.h_min, r.Ksat, r.theta_wp};
        key.insert(key.end(), f, f + 10);
      }
    }
    auto it = seen.find(key);
    if (it == seen.end()) {
      const int id = (int)classes.size();
      classes.push_back(old.invalid_soil_type ? old : make_class_from_rows(pcc, rows.data(), L, thick.data()));
      it = seen.emplace(std::move(key), id).first;
    }
    class_of[(size_t)c] = it->second;
  }
  // new class table, new theta, storage cache
  e->classes.swap(classes);
  e->class_of.swap(class_of);
  e->d_classes = e->dalloc<LgarColumnClass>(e->classes.size(), false);
  e->d.classes = e->d_classes;
  CUDA_TRY(cudaMemcpyAsync(e->d_classes, e->classes.data(), e->classes.size() * sizeof(LgarColumnClass), cudaMemcpyHostToDevice, e->stream));
  CUDA_TRY(cudaMemcpyAsync(e->d.class_id, e->class_of.data(), (size_t)ncol * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
  for (int i = 0; i < maxnf; i++)
    CUDA_TRY(cudaMemcpyAsync(e->d.theta + (size_t)i * st, theta.data() + (size_t)i * ncol, (size_t)ncol * sizeof(double), cudaMemcpyHostToDevice, e->stream));
  CUDA_TRY(cudaMemcpyAsync(e->d.scal + (size_t)LGAR_S_STORAGE * st, storage.data(), (size_t)ncol * sizeof(double), cudaMemcpyHostToDevice, e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
}

// frozen_factor_hydraulic_conductivity (lgar.cxx:1109-1149) for every column: layer-averaged soil temperature ->
// exp(-10 (273.15 - T)) clamped to [0.05, 1].  Host side with libm like the reference; the cell cursor runs on from
// layer to layer as it does there.
void update_frozen_factors(lgarb_engine* e) {
  if (!e->pc.sft_coupled || !e->ff_dirty) return;
  const std::vector<double>& z = e->pc.soil_z;
  const int ncell = (int)z.size(), L = e->num_layers;
  const int64_t ncol = e->ncol;
  const size_t st = (size_t)e->stride;
  std::vector<double> ff((size_t)(LGAR_MAXL + 1) * st, 1.0);
  for (int64_t col = 0; col < ncol; col++) {
    const LgarColumnClass& cc = e->classes[e->class_of[col]];
    if (cc.invalid_soil_type) continue;
    const double* T = e->soil_temperature.data() + (size_t)col * ncell;
    int c = 0;
    for (int layer = 1; layer <= L; layer++) {
      double layer_temp = 0.0;
      int count = 0;
      while (c < ncell && z[c] <= cc.cum[layer]) {
        layer_temp += T[c];
        c++;
        count++;
      }
      if (count == 0) throw std::runtime_error("soil_z has no temperature cell inside soil layer " + std::to_string(layer));  // assert (count > 0), lgar.cxx:1131
      layer_temp /= count;
      double factor = exp(-10 * (273.15 - layer_temp));
      factor = fmax(fmin(factor, 1.0), 0.05);
      ff[(size_t)layer * st + (size_t)col] = factor;
    }
  }
  CUDA_TRY(cudaMemcpyAsync(e->d_ff, ff.data(), ff.size() * sizeof(double), cudaMemcpyHostToDevice, e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  e->ff_dirty = false;
}

void launch_step(lgarb_engine* e, const double* d_precip, const double* d_pet, const LgarSinks& sinks) {
  LgarDeviceState d = e->d;
  d.precip = d_precip;
  d.pet = d_pet;
  d.paths = e->d_paths;
  cudaEvent_t* ev = nullptr;
  if (e->profiling) {
    if (e->ev_used + 3 > e->ev_pool.size()) {
      if (e->ev_pool.size() >= 3 * 4096) {  // bound the pool: fold finished triples into the totals
        CUDA_TRY(cudaStreamSynchronize(e->stream));
        for (size_t i = 0; i + 2 < e->ev_used; i += 3) {
          float a = 0, b = 0;
          cudaEventElapsedTime(&a, e->ev_pool[i], e->ev_pool[i + 1]);
          cudaEventElapsedTime(&b, e->ev_pool[i + 1], e->ev_pool[i + 2]);
          e->prof_ms_stream += a;
          e->prof_ms_active += b;
        }
        e->ev_used = 0;
      } else {
        for (int i = 0; i < 3; i++) {
          cudaEvent_t x;
          CUDA_TRY(cudaEventCreate(&x));
          e->ev_pool.push_back(x);
        }
      }
    }
    ev = e->ev_pool.data() + e->ev_used;
    e->ev_used += 3;
    e->prof_steps++;
  }
  CUDA_TRY(lgar_launch_step(e->cfg, d, sinks, e->active_blocks, e->force_all_active, e->stream, ev));
  if (e->profiling) {
    CUDA_TRY(lgar_launch_sum_fronts(e->d, e->d_prof_acc, e->stream));
    e->launches += 1;
  }
  e->launches += 2;
  e->steps++;
  e->time_s += e->cfg.forcing_resolution_h * 3600.0;  // bmi_lgar.cxx:359 summed over the sub-cycles
}

void require_init(const lgarb_engine* e) {
  if (!e || !e->initialized) throw std::runtime_error("engine is not initialized");
}

void check_range(const lgarb_engine* e, int64_t col0, int64_t ncol) {
  if (col0 < 0 || ncol < 0 || col0 + ncol > e->ncol) throw std::runtime_error("column range out of bounds");
}

// copy a strided device plane [rows][stride] -> host [ncol][rows]?  No: scalar variables are
// [ncol] planes, returned as they are.
void d2h(lgarb_engine* e, void* dst, const void* src, size_t bytes) {
  CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
}

int var_grid(const std::string& name) {  // bmi_lgar.cxx:1112-1154
  if (name == "soil_storage_model" || name == "soil_num_wetting_fronts") return 0;
  static const char* g1[] = {"precipitation_rate", "precipitation", "potential_evapotranspiration_rate", "potential_evapotranspiration",
                             "actual_evapotranspiration", "surface_runoff", "giuh_runoff", "a", "b", "frac_to_CR", "spf_factor", "a_slow",
                             "b_slow", "frac_slow", "soil_storage", "field_capacity", "ponded_depth_max", "total_discharge", "infiltration",
                             "percolation", "conceptual_reservoir_to_stream_discharge", "mass_balance", "smcmax_1", "smcmin_1",
                             "van_genuchten_m_1", "van_genuchten_alpha_1", "van_genuchten_n_1", "hydraulic_conductivity_1", "smcmax_2",
                             "smcmin_2", "van_genuchten_m_2", "van_genuchten_alpha_2", "van_genuchten_n_2", "hydraulic_conductivity_2"};
  for (const char* s : g1)
    if (name == s) return 1;
  if (name == "soil_depth_layers") return 2;
  if (name == "soil_moisture_wetting_fronts" || name == "soil_depth_wetting_fronts") return 3;
  if (name == "soil_temperature_profile") return 4;
  return -1;
}

// where a gettable variable lives: returns the device pointer of a [ncol] plane, or null for the
// gathered (per-front / per-layer) ones
const void* scalar_plane(lgarb_engine* e, const std::string& name, size_t* itemsize) {
  *itemsize = sizeof(double);
  if (name == "precipitation_rate") return e->d_precip;
  if (name == "potential_evapotranspiration_rate") return e->d_pet;
  int k = out_index(name);
  if (k >= 0) return e->d.out + (size_t)k * e->stride;
  if (name == "soil_num_wetting_fronts") {
    *itemsize = sizeof(int32_t);
    return e->d.nfront;
  }
  return nullptr;
}

void get_value_range(lgarb_engine* e, const std::string& name, int64_t col0, int64_t ncol, void* dest, bool dest_is_device) {
  require_init(e);
  check_range(e, col0, ncol);
  size_t isz;
  const void* plane = scalar_plane(e, name, &isz);
  cudaMemcpyKind kind = dest_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  if (plane) {
    CUDA_TRY(cudaMemcpyAsync(dest, (const char*)plane + (size_t)col0 * isz, (size_t)ncol * isz, kind, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return;
  }
  if (name == "soil_moisture_wetting_fronts" || name == "soil_depth_wetting_fronts") {
    bool depth = name == "soil_depth_wetting_fronts";
    CUDA_TRY(lgar_launch_gather_fronts(e->d, depth ? 0 : 1, depth ? 0.01 : 1.0, e->d_stage, col0, ncol, e->stream));  // bmi_lgar.cxx:836-837
    e->launches++;
    CUDA_TRY(cudaMemcpyAsync(dest, e->d_stage, (size_t)ncol * LGAR_W * sizeof(double), kind, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return;
  }
  if (name == "soil_depth_layers") {
    CUDA_TRY(lgar_launch_gather_layers(e->d, e->d_stage, col0, ncol, e->num_layers, e->stream));
    e->launches++;
    CUDA_TRY(cudaMemcpyAsync(dest, e->d_stage, (size_t)ncol * e->num_layers * sizeof(double), kind, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return;
  }
  {  // calibratable parameters (src/bmi_lgar.cxx:1372-1421): one value per column
    const int si = calib_soil_index(name), gi = calib_glob_index(name);
    if (si >= 0 || gi >= 0) {
      std::vector<double> v((size_t)ncol);
      for (int64_t c = 0; c < ncol; c++)
        v[(size_t)c] = gi >= 0 ? (e->calib_glob_col[gi].empty() ? e->calib_glob[gi] : e->calib_glob_col[gi][(size_t)(col0 + c)])
                               : e->calib_soil[si].empty() ? calib_soil_initial(e, si, col0 + c) : e->calib_soil[si][(size_t)(col0 + c)];
      if (dest_is_device) CUDA_TRY(cudaMemcpy(dest, v.data(), (size_t)ncol * sizeof(double), cudaMemcpyHostToDevice));
      else memcpy(dest, v.data(), (size_t)ncol * sizeof(double));
      return;
    }
  }
  throw std::runtime_error("variable " + name + " does not exist");  // bmi_lgar.cxx:1421-1425
}

void set_value_range(lgarb_engine* e, const std::string& name, int64_t col0, int64_t ncol, const void* src, bool src_is_device) {
  require_init(e);
  check_range(e, col0, ncol);
  double* plane = nullptr;
  if (name == "precipitation_rate") plane = e->d_precip;
  else if (name == "potential_evapotranspiration_rate") plane = e->d_pet;
  else if (name == "soil_temperature_profile") {
    // [ncol][num_cells_temp] K (grid 4, src/bmi_lgar.cxx:1144); without SFT coupling the reference keeps one unused cell
    if (!e->pc.sft_coupled) return;
    const size_t ncell = e->pc.soil_z.size();
    double* dst = e->soil_temperature.data() + (size_t)col0 * ncell;
    if (src_is_device) CUDA_TRY(cudaMemcpy(dst, src, (size_t)ncol * ncell * sizeof(double), cudaMemcpyDeviceToHost));
    else memcpy(dst, src, (size_t)ncol * ncell * sizeof(double));
    e->ff_dirty = true;
    return;
  }
  else if (calib_soil_index(name) >= 0 || calib_glob_index(name) >= 0) {
    // SetValue on a calibratable parameter writes lgar_calib_params (src/bmi_lgar.cxx:1455-1467); the values take
    // effect at the first Update of a run configured with calib_params=true (:187-190) and never afterwards
    std::vector<double> v((size_t)ncol);
    if (src_is_device) CUDA_TRY(cudaMemcpy(v.data(), src, (size_t)ncol * sizeof(double), cudaMemcpyDeviceToHost));
    else memcpy(v.data(), src, (size_t)ncol * sizeof(double));
    const int si = calib_soil_index(name), gi = calib_glob_index(name);
    if (si >= 10) {
      // Reference quirk: GetVarGrid lists the *_2 names twice and the *_3 names never (src/bmi_lgar.cxx:1133-1141), so
      // GetVarNbytes("smcmax_3") is 0 and SetValue copies nothing (:1455-1467): the layer-3 parameters cannot be changed
      // through SetValue.  Same here, so that the same calls give the same run.
      return;
    }
    if (si >= 0) {
      std::vector<double>& a = e->calib_soil[si];
      if (a.empty()) {
        a.resize((size_t)e->ncol);
        for (int64_t c = 0; c < e->ncol; c++) a[(size_t)c] = calib_soil_initial(e, si, c);
      }
      for (int64_t c = 0; c < ncol; c++) a[(size_t)(col0 + c)] = v[(size_t)c];
    } else {
      // reservoir / ponding / field-capacity parameters: per column like the soil parameters (they live in the class row)
      std::vector<double>& a = e->calib_glob_col[gi];
      if (a.empty()) a.assign((size_t)e->ncol, e->calib_glob[gi]);
      for (int64_t c = 0; c < ncol; c++) a[(size_t)(col0 + c)] = v[(size_t)c];
    }
    e->calib_touched = true;
    return;
  }
  else if (var_grid(name) >= 0) throw std::runtime_error("variable " + name + " is not settable in this engine");
  else throw std::runtime_error("variable " + name + " does not exist");
  CUDA_TRY(cudaMemcpyAsync(plane + col0, src, (size_t)ncol * sizeof(double), src_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                           e->stream));
  if (!src_is_device) CUDA_TRY(cudaStreamSynchronize(e->stream));  // the caller may reuse src immediately
}

void copy_str(char* dst, int cap, const std::string& s) {
  if (!dst || cap <= 0) throw std::runtime_error("bad string buffer");
  snprintf(dst, (size_t)cap, "%s", s.c_str());
}

}  // namespace

// CUDA's current device is a property of the calling THREAD: every entry point that touches CUDA makes the engine's device
// current for the duration of the call and puts the caller's back afterwards, so that two engines on two GPUs in one process, or
// a call from another thread than the one that initialised the engine, allocate and launch where the engine lives.
struct LgarbDeviceGuard {
  int prev = -1;
  explicit LgarbDeviceGuard(const lgarb_engine* e) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (e && e->has_device && e->device != prev) cudaSetDevice(e->device);
  }
  ~LgarbDeviceGuard() {
    int now = -1;
    if (prev >= 0 && cudaGetDevice(&now) == cudaSuccess && now != prev) cudaSetDevice(prev);
  }
};

#define LGARB_GUARD(e, ...)                                   \
  try {                                                       \
    LgarbDeviceGuard lgarb_device_guard_(e);                  \
    __VA_ARGS__;                                              \
    return LGARB_SUCCESS;                                     \
  } catch (const std::exception& ex) {                        \
    if (e) const_cast<lgarb_engine*>(e)->last_error = ex.what(); \
    return LGARB_FAILURE;                                     \
  } catch (...) {                                             \
    if (e) const_cast<lgarb_engine*>(e)->last_error = "unknown error"; \
    return LGARB_FAILURE;                                     \
  }

extern "C" {

namespace {
void run_wave_window(lgarb_engine* e, LgarWindow w, int use_slots);  // one window through the stage kernels (defined below)
}

int lgarb_create(lgarb_engine** out) {
  if (!out) return LGARB_FAILURE;
  *out = new (std::nothrow) lgarb_engine();
  if (!*out) return LGARB_FAILURE;
  // scheduler tuning knobs for experiments (defaults are the measured optimum, profiles/r1_ncu_summary.md)
  auto knob = [](const char* name, int& v) {
    const char* x = getenv(name);
    if (x && *x) v = atoi(x);
  };
  knob("LGARB_WAVE_QUANTUM", (*out)->wave_quantum);
  knob("LGARB_WAVE_PAIRS", (*out)->wave_pairs);
  knob("LGARB_WAVE_CORR_PAIRS", (*out)->wave_corr_pairs);
  knob("LGARB_WAVE_CORR_ASYNC", (*out)->wave_corr_async);
  knob("LGARB_WAVE_FINISH_BELOW", (*out)->wave_finish_below);
  knob("LGARB_WAVE_FF_BELOW", (*out)->wave_ff_below);
  knob("LGARB_FINISH_SMEM", (*out)->finish_smem);
  knob("LGARB_FINISH_FF", (*out)->finish_ff);
  knob("LGARB_FINISH_SPW", (*out)->finish_spw);
  knob("LGARB_FINISH_QUANTUM", (*out)->finish_quantum);
  knob("LGARB_FINISH_LANE_STRIDE", (*out)->finish_lane_stride);
  knob("LGARB_UPDATE_WAVE_MIN_COLS", (*out)->update_wave_min_cols);
  knob("LGARB_UPDATE_WAVE_SLOTS", (*out)->update_wave_slots);
  knob("LGARB_UPDATE_WAVE_FINISH_BELOW", (*out)->update_wave_finish_below);
  knob("LGARB_HOST_CHUNK_HOURS", (*out)->host_chunk_hours);
  knob("LGARB_WAVE_ESCALATE_BELOW", (*out)->wave_escalate_below);
  knob("LGARB_WAVE_CORR_QUANTUM", (*out)->wave_corr_quantum);
  knob("LGARB_WINDOW_MODE", (*out)->use_window);
  return LGARB_SUCCESS;
}

int lgarb_destroy(lgarb_engine* e) {
  if (!e) return LGARB_FAILURE;
  {
    LgarbDeviceGuard guard(e);
    e->release();  // unconditionally: also what a failed initialisation may have left
  }
  delete e;
  return LGARB_SUCCESS;
}

int lgarb_initialize(lgarb_engine* e, const char* config_file, int64_t n_columns, int device) {
  if (!e || !config_file) return LGARB_FAILURE;
  LGARB_GUARD(e, do_initialize(e, config_file, n_columns, device, nullptr, nullptr));
}

int lgarb_initialize_columns(lgarb_engine* e, const char* config_file, int64_t n_columns, int device, const int32_t* layer_soil_type,
                             const double* layer_thickness_cm) {
  if (!e || !config_file) return LGARB_FAILURE;
  LGARB_GUARD(e, do_initialize(e, config_file, n_columns, device, layer_soil_type, layer_thickness_cm));
}

int lgarb_finalize(lgarb_engine* e) {
  if (!e) return LGARB_FAILURE;
  LGARB_GUARD(e, { require_init(e); e->release(); });
}

const char* lgarb_last_error(const lgarb_engine* e) { return e ? e->last_error.c_str() : "null engine"; }

int lgarb_update(lgarb_engine* e) {
  LGARB_GUARD(e, {
    require_init(e);
    apply_calibration(e);
    update_frozen_factors(e);
    if (e->use_window >= 3 && e->ncol >= e->update_wave_min_cols) {
      // Large batch: the stream kernel finishes the columns that replay cached fluxes (84 % of the column-steps of the bench
      // workload) from their scalars and lists the others; the listed columns then take the stage kernels as a window of one
      // hour -- slots import a column of the list, run its step through HEAD / FRONT / SEARCH / TAIL and export it.  The
      // active kernel of the pair below makes every listed column wait for the slowest psi search of its warp; the stage
      // kernels do not.  (A window of one hour over ALL columns was measured in round 1: the import/export of every column costs
      // more than it saves.)
      LgarDeviceState d = e->d;
      d.precip = e->d_precip;
      d.pet = e->d_pet;
      d.paths = e->d_paths;
      LgarSinks sinks;
      memset(&sinks, 0, sizeof(sinks));
      CUDA_TRY(lgar_launch_step_stream(e->cfg, d, sinks, e->force_all_active, e->stream));
      e->launches += 1;
      LgarWindow w;
      memset(&w, 0, sizeof(w));
      w.precip = e->d_precip;
      w.pet = e->d_pet;
      w.series_stride = e->stride;
      w.sink_stride = e->stride;
      w.K = 1;
      w.force_all_active = e->force_all_active;
      w.col_list = e->d.work;
      w.col_count = e->d.work_count;
      run_wave_window(e, w, e->update_wave_slots);
      e->steps++;
      e->time_s += e->cfg.forcing_resolution_h * 3600.0;
    } else {
      LgarSinks sinks;
      memset(&sinks, 0, sizeof(sinks));
      launch_step(e, e->d_precip, e->d_pet, sinks);
    }
    CUDA_TRY(cudaStreamSynchronize(e->stream));
  });
}

int lgarb_update_until(lgarb_engine* e, double t) {
  if (e && !(t > 0.0)) {  // bmi_lgar.cxx:901 assert (t > 0.0)
    e->last_error = "UpdateUntil: t must be > 0";
    return LGARB_FAILURE;
  }
  return lgarb_update(e);
}

namespace {
void ensure_wave(lgarb_engine* e) {
  if (e->wave_ready) return;
  LgarWave& wv = e->wave;
  memset(&wv, 0, sizeof(wv));
  wv.nslots = (int32_t)std::min<int64_t>(e->ncol, e->wave_max_slots);
  const size_t mark = e->allocs.size();
  try {
    void* p = nullptr;
    CUDA_TRY(cudaMalloc(&p, (size_t)wv.nslots * lgar_lane_context_bytes()));
    e->allocs.push_back(p);
    wv.ctx = (LgLane*)p;
    for (int s = 0; s < LGAR_NSTAGE; s++)
      for (int b = 0; b < 2; b++) wv.queue[s][b] = e->dalloc<int32_t>((size_t)wv.nslots * LGAR_NBIN, false);
    wv.qcount = e->dalloc<int32_t>(LGAR_NSTAGE * 2 * LGAR_NBIN);
    wv.done_columns = e->dalloc<int32_t>(1);
    wv.cursor = e->dalloc<int32_t>(1);
    wv.waiting = e->dalloc<int32_t>(1);
    wv.done_blocks = e->dalloc<int32_t>(LGAR_NSTAGE);
    if (!e->side_stream) {
      CUDA_TRY(cudaStreamCreateWithFlags(&e->side_stream, cudaStreamNonBlocking));
      for (int i = 0; i < 2; i++) CUDA_TRY(cudaEventCreateWithFlags(&e->side_ev[i], cudaEventDisableTiming));
    }
  } catch (...) {  // a retry must not pile up what this attempt allocated
    while (e->allocs.size() > mark) {
      cudaFree(e->allocs.back());
      e->allocs.pop_back();
    }
    memset(&wv, 0, sizeof(wv));
    throw;
  }
  e->wave_ready = true;
}

// one window of K hours through the wavefront kernels
// use_slots > 0: that many slots draw their columns from w.col_list (or 0 .. ncol - 1) instead of slot i owning column i
void run_wave_window(lgarb_engine* e, LgarWindow w, int use_slots) {
  ensure_wave(e);
  LgarWave wv = e->wave;
  if (use_slots > 0) wv.nslots = std::min(wv.nslots, use_slots);
  w.tile_counter = e->d_tile_counter;
  w.done_columns = wv.done_columns;
  w.paths = e->d_paths;
  CUDA_TRY(cudaMemsetAsync(e->d_tile_counter, 0, sizeof(int32_t), e->stream));
  w.slot_is_column = (use_slots <= 0 && wv.nslots == e->ncol) ? 1 : 0;
  w.waiting = wv.waiting;
  lgarb_engine::Frontier* fr = e->frontier;
  int hours_present = fr ? 0 : w.K;
  // raise the frontier to what has arrived; `block`: wait for the next chunk first (nothing can move without it)
  auto advance_frontier = [&](bool block) {
    if (!fr) return;
    if (block && fr->next < fr->ev.size()) {
      CUDA_TRY(cudaEventSynchronize(fr->ev[fr->next]));
      e->frontier_waits++;
    }
    while (fr->next < fr->ev.size()) {
      const cudaError_t q = cudaEventQuery(fr->ev[fr->next]);
      if (q == cudaErrorNotReady) break;
      CUDA_TRY(q);
      hours_present = fr->upto[fr->next++];
    }
    if (fr->next == fr->ev.size()) hours_present = w.K;
    w.frontier1 = hours_present >= w.K ? 0 : hours_present + 1;
  };
  advance_frontier(true);  // the first chunk: nothing to do before it is there
  CUDA_TRY(lgar_launch_wave_init(e->cfg, e->d, wv, w.slot_is_column, e->stream));
  e->launches += 1;
  int par[LGAR_NSTAGE] = {0, 0, 0, 0, 0, 0};
  int quantum = e->wave_quantum;
  int in_flight = wv.nslots;
  const bool dbg_on = getenv("LGARB_WAVE_DEBUG") != nullptr;
  // `ncu --profile-from-start off`: LGARB_NCU_WINDOW=i LGARB_NCU_ROUND=r profiles exactly round r of the i-th window of this engine
  const int ncu_round = (getenv("LGARB_NCU_WINDOW") && atoi(getenv("LGARB_NCU_WINDOW")) == e->wave_windows)
                            ? (getenv("LGARB_NCU_ROUND") ? atoi(getenv("LGARB_NCU_ROUND")) : 0) : -1;
  e->wave_windows++;
  const bool ncu_whole_window = ncu_round == -2;  // LGARB_NCU_ROUND=-2: every launch of the window
  if (ncu_whole_window) {
    cudaStreamSynchronize(e->stream);
    cudaProfilerStart();
  }
  struct NcuStop {
    bool on;
    cudaStream_t s;
    ~NcuStop() {
      if (on) {
        cudaStreamSynchronize(s);
        cudaProfilerStop();
      }
    }
  } ncu_stop{ncu_whole_window, e->stream};
  std::vector<std::pair<cudaEvent_t, int>> dbg_evs;
  cudaStream_t launch_stream = e->stream;
  auto launch = [&](int stage) {
    const int consume = par[stage];
    par[stage] ^= 1;
    unsigned mask = 0;
    for (int s = 0; s < LGAR_NSTAGE; s++) mask |= (unsigned)par[s] << s;
    if (dbg_on) {
      dbg_evs.emplace_back();
      cudaEventCreate(&dbg_evs.back().first);
      cudaEventRecord(dbg_evs.back().first, e->stream);
      dbg_evs.back().second = stage;
    }
    lgarb_engine::StageEv se{nullptr, nullptr, stage};
    if (e->stage_timing) {
      CUDA_TRY(cudaEventCreate(&se.a));
      CUDA_TRY(cudaEventCreate(&se.b));
      CUDA_TRY(cudaEventRecord(se.a, launch_stream));
    }
    CUDA_TRY(lgar_launch_wave_stage(e->cfg, e->d, w, wv, stage, consume, mask, stage == LG_ST_CORR ? std::max(1, quantum * e->wave_corr_quantum / e->wave_quantum) : quantum, in_flight,
                                    e->sm_count, launch_stream));
    if (e->stage_timing) {
      CUDA_TRY(cudaEventRecord(se.b, launch_stream));
      e->stage_evs.push_back(se);
    }
    e->launches += 1;
  };
  int32_t* h_cnt = (int32_t*)e->pinned(64 * sizeof(int32_t));
  const bool debug = getenv("LGARB_WAVE_DEBUG") != nullptr;
  cudaEvent_t dbg[2];
  if (debug) {
    cudaEventCreate(&dbg[0]);
    cudaEventCreate(&dbg[1]);
  }
  // Every round starts with the queue counts on the host (48 bytes): the number of slots in flight bounds
  // every queue of the round (grid size), empty parts of the round are not launched, and the window is over
  // when nothing is in flight.
  int cnt[LGAR_NSTAGE] = {wv.nslots, 0, 0, 0, 0, 0};
  for (int64_t round = 0;; round++) {
    if (round > 0) {
      CUDA_TRY(lgar_launch_wave_poll(wv, h_cnt, e->stream));
      CUDA_TRY(cudaStreamSynchronize(e->stream));
      in_flight = 0;
      for (int s = 0; s < LGAR_NSTAGE; s++) {
        cnt[s] = h_cnt[2 * s] + h_cnt[2 * s + 1];
        in_flight += cnt[s];
      }
      if (in_flight == 0) break;
      // every slot in flight is waiting for forcing: block until the next chunk is there
      advance_frontier(hours_present < w.K && h_cnt[2 * LGAR_NSTAGE] >= in_flight);
      if (in_flight <= (use_slots > 0 ? std::max(e->wave_finish_below, e->update_wave_finish_below) : e->wave_finish_below) && hours_present >= w.K) {  // the stragglers: one thread per slot runs its column to the end of the window
        if (debug) cudaEventRecord(dbg[0], e->stream);
        w.use_ff = e->finish_ff;
        lgarb_engine::StageEv se{nullptr, nullptr, LGAR_NSTAGE};
        if (e->stage_timing) {
          CUDA_TRY(cudaEventCreate(&se.a));
          CUDA_TRY(cudaEventCreate(&se.b));
          CUDA_TRY(cudaEventRecord(se.a, e->stream));
        }
        if (e->finish_smem)
          CUDA_TRY(lgar_launch_wave_finish_smem(e->cfg, e->d, w, wv, in_flight, e->finish_spw, e->finish_quantum, e->finish_ff, e->sm_count, e->stream));
        else
          CUDA_TRY(lgar_launch_wave_finish(e->cfg, e->d, w, wv, in_flight, e->finish_lane_stride, e->finish_ff, e->sm_count, e->stream));
        if (e->stage_timing) {
          CUDA_TRY(cudaEventRecord(se.b, e->stream));
          e->stage_evs.push_back(se);
        }
        e->launches += 2;
        if (debug) {
          cudaEventRecord(dbg[1], e->stream);
          cudaEventSynchronize(dbg[1]);
          float ms = 0;
          cudaEventElapsedTime(&ms, dbg[0], dbg[1]);
          fprintf(stderr, "[wave] round %lld: finishing kernel for %d slots in flight: %.3f ms\n", (long long)round, in_flight, ms);
        }
        break;
      }
      if (round > 100000000) throw std::runtime_error("wavefront scheduler did not terminate");
    }
    // few slots in flight = the long searches (up to 1e5 trips) are what is left: let them run longer per launch
    // (while thousands of other slots are in flight a long search must not hold the round up: it keeps its place in
    // the SEARCH/CORR queue and continues next round)
    quantum = in_flight > e->wave_escalate_below ? e->wave_quantum : in_flight > 1024 ? 32 * e->wave_quantum : 256 * e->wave_quantum;
    if (debug) cudaEventRecord(dbg[0], e->stream);
    if (round == ncu_round) {
      cudaStreamSynchronize(e->stream);
      cudaProfilerStart();
    }
    w.use_ff = in_flight <= e->wave_ff_below ? 1 : 0;  // the lambdas above launch with `w` as it is now
    const bool new_steps = cnt[LG_ST_FETCH] + cnt[LG_ST_HEAD] > 0;
    const bool corr_async = e->use_window == 3 && e->wave_corr_async && !dbg_on;
    bool corr_forked = false;
    if (corr_async && cnt[LG_ST_CORR] > 0) {
      // CORR consumes what TAIL queued in the previous round and appends to the TAIL queue; nothing else on the main
      // stream touches the CORR queues before the join, and appends to the TAIL queue are atomic
      CUDA_TRY(cudaEventRecord(e->side_ev[0], e->stream));
      CUDA_TRY(cudaStreamWaitEvent(e->side_stream, e->side_ev[0], 0));
      launch_stream = e->side_stream;
      launch(LG_ST_CORR);
      launch_stream = e->stream;
      CUDA_TRY(cudaEventRecord(e->side_ev[1], e->side_stream));
      corr_forked = true;
    }
    if (new_steps) {
      launch(LG_ST_FETCH);
      launch(LG_ST_HEAD);  // runs the first stretch of the front loop too
    }
    if (new_steps || cnt[LG_ST_FRONT] + cnt[LG_ST_SEARCH] > 0) {
      if (cnt[LG_ST_FRONT] > 0) launch(LG_ST_FRONT);
      for (int k = 0; k < e->wave_pairs; k++) {
        launch(LG_ST_SEARCH);
        launch(LG_ST_FRONT);
      }
    }
    {
      if (corr_forked) CUDA_TRY(cudaStreamWaitEvent(e->stream, e->side_ev[1], 0));
      launch(LG_ST_TAIL);
      for (int k = 0; k < (corr_async ? 0 : e->wave_corr_pairs); k++) {
        launch(LG_ST_CORR);
        launch(LG_ST_TAIL);
      }
    }
    e->wave_rounds++;
    if (round == ncu_round) {
      cudaStreamSynchronize(e->stream);
      cudaProfilerStop();
    }
    if (debug) {
      cudaEventRecord(dbg[1], e->stream);
      cudaEventSynchronize(dbg[1]);
      float ms = 0;
      cudaEventElapsedTime(&ms, dbg[0], dbg[1]);
      float st_ms[LGAR_NSTAGE] = {0, 0, 0, 0, 0, 0};
      for (size_t i = 0; i < dbg_evs.size(); i++) {
        float x = 0;
        cudaEventElapsedTime(&x, dbg_evs[i].first, i + 1 < dbg_evs.size() ? dbg_evs[i + 1].first : dbg[1]);
        st_ms[dbg_evs[i].second] += x;
      }
      for (auto& pe : dbg_evs) cudaEventDestroy(pe.first);
      dbg_evs.clear();
      fprintf(stderr, "[wave] round %lld: %.3f ms  in flight %d  queues F=%d H=%d FR=%d S=%d T=%d C=%d  ms: FETCH %.2f HEAD %.2f FRONT %.2f SEARCH %.2f TAIL %.2f CORR %.2f\n",
              (long long)round, ms, in_flight, cnt[0], cnt[1], cnt[2], cnt[3], cnt[4], cnt[5], st_ms[0], st_ms[1], st_ms[2], st_ms[3], st_ms[4], st_ms[5]);
    }
  }
  if (debug) {
    cudaEventDestroy(dbg[0]);
    cudaEventDestroy(dbg[1]);
  }
}
}  // namespace

int lgarb_update_steps_device(lgarb_engine* e, int64_t n_steps, const double* d_precip, const double* d_pet, int64_t series_stride,
                              double* const* d_sinks, int64_t sink_stride, int32_t* d_nfronts_sink) {
  LGARB_GUARD(e, {
    require_init(e);
    apply_calibration(e);
    update_frozen_factors(e);  // an SFT-coupled run keeps the current temperature profile for the whole call
    if (n_steps < 0 || !d_precip || !d_pet || series_stride < e->ncol) throw std::runtime_error("bad forcing series");
    if ((d_sinks || d_nfronts_sink) && sink_stride < e->ncol) throw std::runtime_error("bad sink stride");
    if (!e->use_window || (e->pc.sft_coupled && e->use_window != 3)) {  // SFT-coupled runs: hour by hour, or the stage kernels (frozen factors per lane record)
      for (int64_t t = 0; t < n_steps; t++) {
        LgarSinks sinks;
        for (int k = 0; k < LGAR_NOUT; k++) sinks.p[k] = (d_sinks && d_sinks[k]) ? d_sinks[k] + t * sink_stride : nullptr;
        launch_step(e, d_precip + t * series_stride, d_pet + t * series_stride, sinks);
        if (d_nfronts_sink)
          CUDA_TRY(cudaMemcpyAsync(d_nfronts_sink + t * sink_stride, e->d.nfront, (size_t)e->ncol * sizeof(int32_t), cudaMemcpyDeviceToDevice,
                                   e->stream));
      }
    } else {
      for (int64_t t0 = 0; t0 < n_steps; t0 += LGAR_MAX_WINDOW) {
        const int K = (int)std::min<int64_t>(LGAR_MAX_WINDOW, n_steps - t0);
        LgarWindow w;
        memset(&w, 0, sizeof(w));
        w.precip = d_precip + t0 * series_stride;
        w.pet = d_pet + t0 * series_stride;
        w.series_stride = series_stride;
        for (int k = 0; k < LGAR_NOUT; k++) w.sinks.p[k] = (d_sinks && d_sinks[k]) ? d_sinks[k] + t0 * sink_stride : nullptr;
        w.sink_stride = sink_stride;
        w.has_sinks = 0;
        for (int k = 0; k < LGAR_NOUT; k++) w.has_sinks |= (w.sinks.p[k] != nullptr);
        w.nf_sink = d_nfronts_sink ? d_nfronts_sink + t0 * sink_stride : nullptr;
        w.K = K;
        w.force_all_active = e->force_all_active;
        w.tile_counter = e->d_tile_counter;
        w.prof = e->profiling ? e->d_prof_acc : nullptr;
        w.work_trace = e->d_work_trace;
        if (e->d_work_trace) CUDA_TRY(cudaMemsetAsync(e->d_work_trace + 2 * e->stride, 0, (64 + 2 * 8192) * sizeof(unsigned int), e->stream));
        w.trace_stride = e->stride;
        if (e->profiling) {
          if (!e->win_ev[0]) {
            CUDA_TRY(cudaEventCreate(&e->win_ev[0]));
            CUDA_TRY(cudaEventCreate(&e->win_ev[1]));
          } else if (e->prof_windows > 0) {  // fold the previous window's time before reusing the events
            CUDA_TRY(cudaEventSynchronize(e->win_ev[1]));
            float ms = 0;
            CUDA_TRY(cudaEventElapsedTime(&ms, e->win_ev[0], e->win_ev[1]));
            e->prof_ms_window += ms;
          }
          CUDA_TRY(cudaEventRecord(e->win_ev[0], e->stream));
        }
        if (e->use_window >= 3)
          run_wave_window(e, w, 0);
        else
          CUDA_TRY(lgar_launch_window(e->cfg, e->d, w, e->window_blocks, e->stream, e->use_window));
        if (e->profiling) {
          CUDA_TRY(cudaEventRecord(e->win_ev[1], e->stream));
          e->prof_windows++;
          e->prof_steps += K;
        }
        e->launches += 1;
        e->steps += K;
        e->time_s += K * e->cfg.forcing_resolution_h * 3600.0;
      }
    }
  });
}

// Host-buffer variant of the multi-step call: forcing in, chosen outputs out, copies included.
int lgarb_update_steps_host(lgarb_engine* e, int64_t n_steps, const double* h_precip, const double* h_pet, int n_outputs,
                            const char* const* output_names, double* const* h_outputs) {
  return lgarb_update_steps_host_pitched(e, n_steps, h_precip, h_pet, e ? e->ncol : 0, n_outputs, output_names, h_outputs, e ? e->ncol : 0);
}

// The same with host arrays that are wider than this engine's batch: row t of the forcing starts at h_precip + t * forcing_pitch,
// row t of output i at h_outputs[i] + t * output_pitch (elements).  One engine of a multi-device group works on its block of
// columns of global [step][all columns] arrays this way (lgar_group.cu).
int lgarb_update_steps_host_pitched(lgarb_engine* e, int64_t n_steps, const double* h_precip, const double* h_pet, int64_t forcing_pitch, int n_outputs,
                                    const char* const* output_names, double* const* h_outputs, int64_t output_pitch) {
  LGARB_GUARD(e, {
    require_init(e);
    if (n_steps <= 0 || !h_precip || !h_pet) throw std::runtime_error("bad forcing series");
    if (forcing_pitch < e->ncol || (n_outputs > 0 && output_pitch < e->ncol)) throw std::runtime_error("host pitch smaller than the number of columns");
    const size_t row = (size_t)e->ncol * sizeof(double);
    auto h2d_rows = [&](double* dst, const double* src, int64_t t0, int64_t t1, cudaStream_t st) {
      if (forcing_pitch == e->ncol)
        CUDA_TRY(cudaMemcpyAsync(dst + (size_t)t0 * e->ncol, src + (size_t)t0 * e->ncol, (size_t)(t1 - t0) * row, cudaMemcpyHostToDevice, st));
      else
        CUDA_TRY(cudaMemcpy2DAsync(dst + (size_t)t0 * e->ncol, row, src + (size_t)t0 * forcing_pitch, (size_t)forcing_pitch * sizeof(double), row, (size_t)(t1 - t0),
                                   cudaMemcpyHostToDevice, st));
    };
    if (n_outputs < 0 || n_outputs > LGAR_NOUT) throw std::runtime_error("bad output count");
    const size_t n = (size_t)n_steps * (size_t)e->ncol;
    const size_t need = n * (2 + (size_t)n_outputs);
    if (need > e->host_series_cap) {  // device staging for forcing + requested output series, grown on demand
      if (e->d_host_series) {
        CUDA_TRY(cudaStreamSynchronize(e->stream));
        cudaFree(e->d_host_series);
        e->d_host_series = nullptr;
        e->host_series_cap = 0;
      }
      CUDA_TRY(cudaMalloc((void**)&e->d_host_series, need * sizeof(double)));
      e->host_series_cap = need;
    }
    double* dP = e->d_host_series;
    double* dE = dP + n;
    // The forcing travels in chunks of host_chunk_hours on a stream of its own while the window already runs on what has arrived
    // (the wavefront scheduler walks a column no further than the frontier, run_wave_window): the copy of all but the first
    // chunk hides behind the computation.  Hour-by-hour and single-window-kernel modes copy everything first.
    lgarb_engine::Frontier fr;
    struct FrontierGuard {
      lgarb_engine* e;
      lgarb_engine::Frontier* f;
      ~FrontierGuard() {
        e->frontier = nullptr;
        for (cudaEvent_t x : f->ev) cudaEventDestroy(x);
      }
    } frontier_guard{e, &fr};
    const int64_t ch = e->host_chunk_hours;
    if (e->use_window == 3 && ch > 0 && n_steps > ch && n_steps <= LGAR_MAX_WINDOW) {
      CUDA_TRY(cudaStreamSynchronize(e->stream));  // the staging buffers may still be read by what the caller enqueued before
      if (!e->h2d_stream) CUDA_TRY(cudaStreamCreateWithFlags(&e->h2d_stream, cudaStreamNonBlocking));
      for (int64_t t0 = 0; t0 < n_steps; t0 += ch) {
        const int64_t t1 = std::min(n_steps, t0 + ch);
        h2d_rows(dP, h_precip, t0, t1, e->h2d_stream);
        h2d_rows(dE, h_pet, t0, t1, e->h2d_stream);
        cudaEvent_t ev;
        CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        fr.ev.push_back(ev);
        fr.upto.push_back((int)t1);
        CUDA_TRY(cudaEventRecord(ev, e->h2d_stream));
      }
      e->frontier = &fr;
    } else {
      h2d_rows(dP, h_precip, 0, n_steps, e->stream);
      h2d_rows(dE, h_pet, 0, n_steps, e->stream);
    }
    double* sinks[LGAR_NOUT];
    for (int k = 0; k < LGAR_NOUT; k++) sinks[k] = nullptr;
    for (int i = 0; i < n_outputs; i++) {
      int k = out_index(output_names[i]);
      if (k < 0) throw std::runtime_error(std::string("variable ") + output_names[i] + " is not a per-step scalar output");
      sinks[k] = dE + n * (size_t)(1 + i);
    }
    if (lgarb_update_steps_device(e, n_steps, dP, dE, e->ncol, sinks, e->ncol, nullptr) != LGARB_SUCCESS) throw std::runtime_error(e->last_error);
    for (int i = 0; i < n_outputs; i++) {
      if (output_pitch == e->ncol)
        CUDA_TRY(cudaMemcpyAsync(h_outputs[i], dE + n * (size_t)(1 + i), n * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
      else
        CUDA_TRY(cudaMemcpy2DAsync(h_outputs[i], (size_t)output_pitch * sizeof(double), dE + n * (size_t)(1 + i), row, row, (size_t)n_steps, cudaMemcpyDeviceToHost,
                                   e->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(e->stream));
  });
}

// ---- streamed multi-step run (forcing / output streaming) ----------------------
namespace {
__global__ void lgarb_f32_to_f64_kernel(const float* __restrict__ src, double* __restrict__ dst, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = (double)src[i];
}
__global__ void lgarb_f64_to_f32_kernel(const double* __restrict__ src, float* __restrict__ dst, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = (float)src[i];
}
}  // namespace
extern "C++" {
struct StreamBufs {
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
  cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_d2h[2] = {nullptr, nullptr};
  void* raw_in[2] = {nullptr, nullptr};     // f32 forcing as delivered (only for f32 input)
  double* forcing[2] = {nullptr, nullptr};  // [2][chunk][ncol] f64: precip plane, pet plane
  double* out[2] = {nullptr, nullptr};      // [n_out][chunk][ncol] f64
  float* out32[2] = {nullptr, nullptr};
  size_t cap_forcing[2] = {0, 0}, cap_raw[2] = {0, 0}, cap_out[2] = {0, 0}, cap_out32[2] = {0, 0};  // bytes
  void free_all() {
    for (int b = 0; b < 2; b++) {
      if (raw_in[b]) cudaFree(raw_in[b]);
      if (forcing[b]) cudaFree(forcing[b]);
      if (out[b]) cudaFree(out[b]);
      if (out32[b]) cudaFree(out32[b]);
      if (ev_h2d[b]) cudaEventDestroy(ev_h2d[b]);
      if (ev_comp[b]) cudaEventDestroy(ev_comp[b]);
      if (ev_d2h[b]) cudaEventDestroy(ev_d2h[b]);
      raw_in[b] = nullptr; forcing[b] = nullptr; out[b] = nullptr; out32[b] = nullptr;
      ev_h2d[b] = ev_comp[b] = ev_d2h[b] = nullptr;
      cap_forcing[b] = cap_raw[b] = cap_out[b] = cap_out32[b] = 0;
    }
    if (s_h2d) cudaStreamDestroy(s_h2d);
    if (s_d2h) cudaStreamDestroy(s_d2h);
    s_h2d = s_d2h = nullptr;
  }
};
void free_stream_bufs(StreamBufs* p) {
  if (p) {
    p->free_all();
    delete p;
  }
}
}  // extern "C++"
namespace {
}  // namespace

// n_steps forcing hours from host memory in chunks of chunk_steps: the host-to-device copy of chunk k+1 and the
// device-to-host copy of the outputs of chunk k-1 run on their own streams while chunk k computes (two staging
// buffers each way).  Forcing and outputs may be f64 or f32 on the host side; the engine computes in f64.
#define STREAM_TRY(x)                          \
  do {                                         \
    cudaError_t err__ = (x);                   \
    if (err__ != cudaSuccess) fail(#x, err__); \
  } while (0)
int lgarb_update_steps_host_stream(lgarb_engine* e, int64_t n_steps, int64_t chunk_steps, int forcing_dtype, const void* h_precip,
                                   const void* h_pet, int n_outputs, const char* const* output_names, int output_dtype, void* const* h_outputs) {
  LGARB_GUARD(e, {
    require_init(e);
    if (n_steps <= 0 || !h_precip || !h_pet) throw std::runtime_error("bad forcing series");
    if (n_outputs < 0 || n_outputs > LGAR_NOUT) throw std::runtime_error("bad output count");
    if ((forcing_dtype != LGARB_F64 && forcing_dtype != LGARB_F32) || (output_dtype != LGARB_F64 && output_dtype != LGARB_F32))
      throw std::runtime_error("dtype must be LGARB_F64 or LGARB_F32");
    if (chunk_steps <= 0 || chunk_steps > n_steps) chunk_steps = n_steps;
    chunk_steps = std::min<int64_t>(chunk_steps, LGAR_MAX_WINDOW);
    const size_t ncol = (size_t)e->ncol, cn = (size_t)chunk_steps * ncol;
    const size_t in_item = forcing_dtype == LGARB_F32 ? 4 : 8, out_item = output_dtype == LGARB_F32 ? 4 : 8;
    int out_k[LGAR_NOUT];
    for (int i = 0; i < n_outputs; i++) {
      out_k[i] = out_index(output_names[i]);
      if (out_k[i] < 0) throw std::runtime_error(std::string("variable ") + output_names[i] + " is not a per-step scalar output");
      if (!h_outputs || !h_outputs[i]) throw std::runtime_error("null output buffer");
    }
    if (!e->sbufs) e->sbufs = new StreamBufs();
    StreamBufs& sb = *e->sbufs;  // streams, events and staging live on between calls (freed by finalize)
    auto fail = [&](const char* what, cudaError_t err) {
      cudaDeviceSynchronize();
      sb.free_all();
      throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(err));
    };
    if (!sb.s_h2d) {
      STREAM_TRY(cudaStreamCreateWithFlags(&sb.s_h2d, cudaStreamNonBlocking));
      STREAM_TRY(cudaStreamCreateWithFlags(&sb.s_d2h, cudaStreamNonBlocking));
      for (int b = 0; b < 2; b++) {
        STREAM_TRY(cudaEventCreateWithFlags(&sb.ev_h2d[b], cudaEventDisableTiming));
        STREAM_TRY(cudaEventCreateWithFlags(&sb.ev_comp[b], cudaEventDisableTiming));
        STREAM_TRY(cudaEventCreateWithFlags(&sb.ev_d2h[b], cudaEventDisableTiming));
      }
    }
    const int nbuf = chunk_steps < n_steps ? 2 : 1;
    auto grow = [&](void** p, size_t& cap, size_t bytes) {
      if (bytes <= cap) return;
      if (*p) {
        STREAM_TRY(cudaDeviceSynchronize());
        cudaFree(*p);
        *p = nullptr;
        cap = 0;
      }
      STREAM_TRY(cudaMalloc(p, bytes));
      cap = bytes;
    };
    for (int b = 0; b < nbuf; b++) {
      grow((void**)&sb.forcing[b], sb.cap_forcing[b], 2 * cn * sizeof(double));
      if (forcing_dtype == LGARB_F32) grow(&sb.raw_in[b], sb.cap_raw[b], 2 * cn * sizeof(float));
      if (n_outputs) {
        grow((void**)&sb.out[b], sb.cap_out[b], (size_t)n_outputs * cn * sizeof(double));
        if (output_dtype == LGARB_F32) grow((void**)&sb.out32[b], sb.cap_out32[b], (size_t)n_outputs * cn * sizeof(float));
      }
    }
    const int64_t nchunks = (n_steps + chunk_steps - 1) / chunk_steps;
    auto chunk_len = [&](int64_t k) { return (size_t)std::min<int64_t>(chunk_steps, n_steps - k * chunk_steps); };
    auto enqueue_h2d = [&](int64_t k) {
      const int b = (int)(k % nbuf);
      const size_t n = chunk_len(k) * ncol, off = (size_t)k * (size_t)chunk_steps * ncol;
      // the buffer was last read by the compute of chunk k - nbuf
      if (k >= nbuf) STREAM_TRY(cudaStreamWaitEvent(sb.s_h2d, sb.ev_comp[b], 0));
      char* dst = forcing_dtype == LGARB_F32 ? (char*)sb.raw_in[b] : (char*)sb.forcing[b];
      STREAM_TRY(cudaMemcpyAsync(dst, (const char*)h_precip + off * in_item, n * in_item, cudaMemcpyHostToDevice, sb.s_h2d));
      STREAM_TRY(cudaMemcpyAsync(dst + cn * in_item, (const char*)h_pet + off * in_item, n * in_item, cudaMemcpyHostToDevice, sb.s_h2d));
      STREAM_TRY(cudaEventRecord(sb.ev_h2d[b], sb.s_h2d));
    };
    enqueue_h2d(0);
    for (int64_t k = 0; k < nchunks; k++) {
      const int b = (int)(k % nbuf);
      const size_t len = chunk_len(k), n = len * ncol;
      if (k + 1 < nchunks && nbuf == 2) enqueue_h2d(k + 1);  // overlaps the compute of chunk k
      STREAM_TRY(cudaStreamWaitEvent(e->stream, sb.ev_h2d[b], 0));
      if (k >= nbuf && n_outputs) STREAM_TRY(cudaStreamWaitEvent(e->stream, sb.ev_d2h[b], 0));  // output buffer b drained
      if (forcing_dtype == LGARB_F32) {
        lgarb_f32_to_f64_kernel<<<1184, 256, 0, e->stream>>>((const float*)sb.raw_in[b], sb.forcing[b], n);
        lgarb_f32_to_f64_kernel<<<1184, 256, 0, e->stream>>>((const float*)sb.raw_in[b] + cn, sb.forcing[b] + cn, n);
        e->launches += 2;
      }
      double* sinks[LGAR_NOUT];
      for (int q = 0; q < LGAR_NOUT; q++) sinks[q] = nullptr;
      for (int i = 0; i < n_outputs; i++) sinks[out_k[i]] = sb.out[b] + (size_t)i * cn;
      if (lgarb_update_steps_device(e, (int64_t)len, sb.forcing[b], sb.forcing[b] + cn, e->ncol, n_outputs ? sinks : nullptr, e->ncol, nullptr) != LGARB_SUCCESS) {
        const std::string msg = e->last_error;
        cudaDeviceSynchronize();
        throw std::runtime_error(msg);
      }
      if (output_dtype == LGARB_F32)
        for (int i = 0; i < n_outputs; i++) {
          lgarb_f64_to_f32_kernel<<<1184, 256, 0, e->stream>>>(sb.out[b] + (size_t)i * cn, sb.out32[b] + (size_t)i * cn, n);
          e->launches += 1;
        }
      STREAM_TRY(cudaEventRecord(sb.ev_comp[b], e->stream));
      if (n_outputs) {
        STREAM_TRY(cudaStreamWaitEvent(sb.s_d2h, sb.ev_comp[b], 0));
        const size_t off = (size_t)k * (size_t)chunk_steps * ncol;
        for (int i = 0; i < n_outputs; i++) {
          const char* src = output_dtype == LGARB_F32 ? (const char*)(sb.out32[b] + (size_t)i * cn) : (const char*)(sb.out[b] + (size_t)i * cn);
          STREAM_TRY(cudaMemcpyAsync((char*)h_outputs[i] + off * out_item, src, n * out_item, cudaMemcpyDeviceToHost, sb.s_d2h));
        }
        STREAM_TRY(cudaEventRecord(sb.ev_d2h[b], sb.s_d2h));
      }
    }
    STREAM_TRY(cudaStreamSynchronize(e->stream));
    STREAM_TRY(cudaStreamSynchronize(sb.s_d2h));
    STREAM_TRY(cudaStreamSynchronize(sb.s_h2d));
  });
}

#undef STREAM_TRY

int lgarb_synchronize(lgarb_engine* e) {
  LGARB_GUARD(e, { require_init(e); CUDA_TRY(cudaStreamSynchronize(e->stream)); });
}

int lgarb_set_value(lgarb_engine* e, const char* name, const void* src) {
  LGARB_GUARD(e, { require_init(e); set_value_range(e, name, 0, e->ncol, src, false); });
}
int lgarb_set_value_range(lgarb_engine* e, const char* name, int64_t col0, int64_t ncol, const void* src) {
  LGARB_GUARD(e, set_value_range(e, name, col0, ncol, src, false));
}
int lgarb_set_value_device(lgarb_engine* e, const char* name, const void* d_src) {
  LGARB_GUARD(e, { require_init(e); set_value_range(e, name, 0, e->ncol, d_src, true); });
}
int lgarb_set_value_at_indices(lgarb_engine* e, const char* name, const int* inds, int count, const void* src) {
  LGARB_GUARD(e, {
    require_init(e);
    for (int i = 0; i < count; i++) set_value_range(e, name, inds[i], 1, (const double*)src + i, false);
  });
}

int lgarb_get_value(lgarb_engine* e, const char* name, void* dest) {
  LGARB_GUARD(e, { require_init(e); get_value_range(e, name, 0, e->ncol, dest, false); });
}
int lgarb_get_value_range(lgarb_engine* e, const char* name, int64_t col0, int64_t ncol, void* dest) {
  LGARB_GUARD(e, get_value_range(e, name, col0, ncol, dest, false));
}
int lgarb_get_value_device(lgarb_engine* e, const char* name, void* d_dest) {
  LGARB_GUARD(e, { require_init(e); get_value_range(e, name, 0, e->ncol, d_dest, true); });
}
int lgarb_get_value_ptr(lgarb_engine* e, const char* name, void** d_ptr) {
  LGARB_GUARD(e, {
    require_init(e);
    size_t isz;
    const void* p = scalar_plane(e, name, &isz);
    if (!p) throw std::runtime_error(std::string("variable ") + name + " has no flat device array");
    *d_ptr = const_cast<void*>(p);
  });
}
int lgarb_get_value_at_indices(lgarb_engine* e, const char* name, void* dest, const int* inds, int count) {
  LGARB_GUARD(e, {
    require_init(e);
    size_t isz;
    const void* plane = scalar_plane(e, name, &isz);
    if (!plane) throw std::runtime_error(std::string("variable ") + name + " is not addressable by column index");
    for (int i = 0; i < count; i++) {
      check_range(e, inds[i], 1);
      CUDA_TRY(cudaMemcpyAsync((char*)dest + (size_t)i * isz, (const char*)plane + (size_t)inds[i] * isz, isz, cudaMemcpyDeviceToHost, e->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(e->stream));
  });
}

int lgarb_get_component_name(const lgarb_engine* e, char* name, int cap) {
  LGARB_GUARD(e, copy_str(name, cap, "LASAM (Lumped Arid/Semi-arid Model)"));
}
int lgarb_get_input_item_count(const lgarb_engine* e, int* count) { LGARB_GUARD(e, *count = 3); }
int lgarb_get_output_item_count(const lgarb_engine* e, int* count) { LGARB_GUARD(e, *count = 15); }
const char* lgarb_get_input_var_name(const lgarb_engine*, int i) { return (i >= 0 && i < 3) ? kInputNames[i] : nullptr; }
const char* lgarb_get_output_var_name(const lgarb_engine*, int i) { return (i >= 0 && i < 15) ? kOutputNames[i] : nullptr; }
int lgarb_get_var_grid(const lgarb_engine* e, const char* name, int* grid) { LGARB_GUARD(e, *grid = var_grid(name)); }
int lgarb_get_var_type(const lgarb_engine* e, const char* name, char* type, int cap) {
  LGARB_GUARD(e, {
    int g = var_grid(name);
    copy_str(type, cap, g == 0 ? "int" : (g >= 1 && g <= 4) ? "double" : "none");
  });
}
int lgarb_get_var_itemsize(const lgarb_engine* e, const char* name, int* size) {
  LGARB_GUARD(e, {
    int g = var_grid(name);
    *size = g == 0 ? (int)sizeof(int) : (g >= 1 && g <= 4) ? (int)sizeof(double) : 0;
  });
}
int lgarb_get_var_units(const lgarb_engine* e, const char* name, char* units, int cap) {
  LGARB_GUARD(e, {
    std::string n = name, u = "none";
    if (n == "precipitation_rate" || n == "potential_evapotranspiration_rate") u = "mm h^-1";
    else if (n == "soil_moisture_wetting_fronts") u = "none";
    else if (n == "soil_temperature_profile") u = "K";
    else if (out_index(n) >= 0 && n != "conceptual_reservoir_storage") u = "m";
    else if (n == "soil_depth_layers" || n == "soil_depth_wetting_fronts") u = "m";
    copy_str(units, cap, u);
  });
}
int lgarb_get_grid_size(const lgarb_engine* e, int grid, int64_t* size) {
  LGARB_GUARD(e, {
    require_init(e);
    if (grid == 0 || grid == 1) *size = e->ncol;
    else if (grid == 2) *size = e->ncol * e->num_layers;
    else if (grid == 3) *size = e->ncol * LGAR_W;  // padded; the reference's is the live front count
    else if (grid == 4) *size = e->pc.sft_coupled ? (int64_t)e->pc.soil_z.size() * e->ncol : e->ncol;  // num_cells_temp per column (1 without SFT, lgar.cxx:956-958)
    else *size = -1;
  });
}
int lgarb_get_var_nbytes(const lgarb_engine* e, const char* name, int64_t* nbytes) {
  LGARB_GUARD(e, {
    int isz = 0;
    int64_t gs = 0;
    int g = var_grid(name);
    if (lgarb_get_var_itemsize(e, name, &isz) != LGARB_SUCCESS || lgarb_get_grid_size(e, g, &gs) != LGARB_SUCCESS)
      throw std::runtime_error(e->last_error);
    *nbytes = (int64_t)isz * gs;
  });
}
int lgarb_get_var_location(const lgarb_engine* e, const char* name, char* loc, int cap) {
  LGARB_GUARD(e, {
    std::string n = name;
    bool node = (var_grid(n) >= 0) && n != "field_capacity" && n != "ponded_depth_max" && n.rfind("smc", 0) != 0 &&
                n.rfind("van_genuchten", 0) != 0 && n.rfind("hydraulic_conductivity", 0) != 0 && n != "soil_storage_model";
    copy_str(loc, cap, node ? "node" : "none");
  });
}
int lgarb_get_current_time(const lgarb_engine* e, double* t) { LGARB_GUARD(e, { require_init(e); *t = e->time_s; }); }
int lgarb_get_start_time(const lgarb_engine* e, double* t) { LGARB_GUARD(e, *t = 0.0); }
int lgarb_get_end_time(const lgarb_engine* e, double* t) { LGARB_GUARD(e, { require_init(e); *t = e->pc.endtime_s; }); }
int lgarb_get_time_step(const lgarb_engine* e, double* dt) {
  LGARB_GUARD(e, { require_init(e); *dt = e->cfg.forcing_resolution_h * 3600.; });
}
int lgarb_get_time_units(const lgarb_engine* e, char* units, int cap) { LGARB_GUARD(e, copy_str(units, cap, "s")); }
int lgarb_get_grid_rank(const lgarb_engine* e, int grid, int* rank) { LGARB_GUARD(e, *rank = (grid >= 0 && grid <= 4) ? 1 : -1); }
int lgarb_get_grid_type(const lgarb_engine* e, int grid, char* type, int cap) {
  LGARB_GUARD(e, copy_str(type, cap, grid == 0 ? "uniform_rectilinear" : ""));
}

int64_t lgarb_get_num_columns(const lgarb_engine* e) { return (e && e->initialized) ? e->ncol : -1; }
int lgarb_get_num_layers(const lgarb_engine* e) { return (e && e->initialized) ? e->num_layers : -1; }
int lgarb_get_num_giuh_ordinates(const lgarb_engine* e) { return (e && e->initialized) ? e->cfg.num_giuh : -1; }
const char* lgarb_output_scalar_name(int k) { return (k >= 0 && k < LGAR_NOUT) ? kOutNames[k] : nullptr; }

int lgarb_get_status(lgarb_engine* e, int64_t col0, int64_t ncol, int32_t* dest) {
  LGARB_GUARD(e, {
    require_init(e);
    check_range(e, col0, ncol);
    d2h(e, dest, e->d.status + col0, (size_t)ncol * sizeof(int32_t));
  });
}

int lgarb_get_fronts(lgarb_engine* e, int64_t col0, int64_t ncol, double* depth_cm, double* theta, double* psi_cm, double* K, double* dzdt,
                     int32_t* layer_num, int32_t* to_bottom, int32_t* nfronts) {
  LGARB_GUARD(e, {
    require_init(e);
    check_range(e, col0, ncol);
    double* dsts[5] = {depth_cm, theta, psi_cm, K, dzdt};
    for (int w = 0; w < 5; w++) {
      if (!dsts[w]) continue;
      CUDA_TRY(lgar_launch_gather_fronts(e->d, w, 1.0, e->d_stage, col0, ncol, e->stream));
      e->launches++;
      d2h(e, dsts[w], e->d_stage, (size_t)ncol * LGAR_W * sizeof(double));
    }
    std::vector<int32_t> nf((size_t)ncol);
    d2h(e, nf.data(), e->d.nfront + col0, (size_t)ncol * sizeof(int32_t));
    if (nfronts) memcpy(nfronts, nf.data(), (size_t)ncol * sizeof(int32_t));
    if (layer_num || to_bottom) {
      std::vector<unsigned long long> fl((size_t)ncol * LGAR_FW);  // [word][column]
      for (int wd = 0; wd < LGAR_FW; wd++)
        d2h(e, fl.data() + (size_t)wd * ncol, e->d.fflags + (size_t)wd * e->stride + col0, (size_t)ncol * sizeof(unsigned long long));
      for (int64_t c = 0; c < ncol; c++)
        for (int i = 0; i < LGAR_W; i++) {
          int nib = (i < nf[c]) ? (int)((fl[(size_t)(i >> 4) * ncol + c] >> (4 * (i & 15))) & 0xF) : 0;
          if (layer_num) layer_num[c * LGAR_W + i] = (i < nf[c]) ? (nib & 7) : 0;
          if (to_bottom) to_bottom[c * LGAR_W + i] = (i < nf[c]) ? (nib >> 3) : 0;
        }
    }
  });
}

int lgarb_get_global_mass_balance(lgarb_engine* e, int64_t col0, int64_t ncol, double* out) {
  LGARB_GUARD(e, {
    require_init(e);
    check_range(e, col0, ncol);
    const size_t n = (size_t)ncol, st = (size_t)e->stride;
    std::vector<double> cum(n * LGAR_NCUM), vs(n), volon(n), volend(n), crf(n), crs(n);
    for (int k = 0; k < LGAR_NCUM; k++) d2h(e, cum.data() + k * n, e->d.cum + k * st + col0, n * sizeof(double));
    d2h(e, vs.data(), e->volstart + col0, n * sizeof(double));
    d2h(e, volon.data(), e->d.scal + LGAR_S_VOLON * st + col0, n * sizeof(double));
    d2h(e, volend.data(), e->d.scal + LGAR_S_STORAGE * st + col0, n * sizeof(double));
    d2h(e, crf.data(), e->d.scal + LGAR_S_CR_FAST * st + col0, n * sizeof(double));
    d2h(e, crs.data(), e->d.scal + LGAR_S_CR_SLOW * st + col0, n * sizeof(double));
    for (size_t c = 0; c < n; c++) {
      double* o = out + c * 16;
      double volCRend = crf[c] + crs[c];
      o[0] = vs[c];
      o[1] = cum[LGAR_CUM_PRECIP * n + c];
      o[2] = cum[LGAR_CUM_IN * n + c];
      o[3] = volend[c];
      o[4] = volon[c];
      o[5] = cum[LGAR_CUM_RUNOFF * n + c];
      o[6] = cum[LGAR_CUM_GIUH * n + c];
      o[7] = cum[LGAR_CUM_RECH * n + c];
      o[8] = cum[LGAR_CUM_AET * n + c];
      o[9] = cum[LGAR_CUM_PET * n + c];
      o[10] = cum[LGAR_CUM_RUNOFF_CR * n + c];
      o[11] = volCRend;
      o[12] = cum[LGAR_CUM_Q * n + c];
      o[13] = e->volchange_calib.empty() ? 0.0 : e->volchange_calib[(size_t)col0 + c];
      o[14] = cum[LGAR_CUM_Q_CR * n + c];
      o[15] = o[0] + o[1] - o[5] - o[8] - o[4] - o[7] - o[3] + o[13] - o[10] - o[11];  // lgar.cxx:1185
    }
  });
}

// The same 16 terms for every column of the engine, computed and left ON THE DEVICE ([n_columns][16], cm) behind the engine's
// stream: what a multi-device gather sends (lgar_group.cu).  Arithmetic and order are those of the host version above.
int lgarb_get_global_mass_balance_device(lgarb_engine* e, double* d_out) {
  LGARB_GUARD(e, {
    require_init(e);
    if (!d_out) throw std::runtime_error("null destination");
    double* d_vc = nullptr;
    if (!e->volchange_calib.empty()) {
      d_vc = e->d_stage;  // [stride] staging of the engine
      CUDA_TRY(cudaMemcpyAsync(d_vc, e->volchange_calib.data(), (size_t)e->ncol * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    }
    CUDA_TRY(lgar_launch_mass_balance(e->d, e->volstart, d_vc, d_out, e->stream));
    e->launches += 1;
  });
}

int lgarb_get_soil_row(const lgarb_engine* e, int soil, double* o) {
  LGARB_GUARD(e, {
    require_init(e);
    if (soil < 1 || soil >= (int)e->soils.size()) throw std::runtime_error("soil index out of range");
    const SoilRow& s = e->soils[soil];
    o[0] = s.theta_r; o[1] = s.theta_e; o[2] = s.alpha; o[3] = s.n; o[4] = s.m; o[5] = s.lambda; o[6] = s.psib; o[7] = s.h_min;
    o[8] = s.Ksat; o[9] = s.theta_wp;
  });
}

int lgarb_set_force_all_active(lgarb_engine* e, int on) { LGARB_GUARD(e, e->force_all_active = on ? 1 : 0); }
int lgarb_set_work_trace(lgarb_engine* e, int on) {
  LGARB_GUARD(e, {
    require_init(e);
    if (on && !e->d_work_trace) e->d_work_trace = e->dalloc<unsigned int>(2 * (size_t)e->stride + 64 + 2 * 8192);
    if (!on) e->d_work_trace = nullptr;  // the buffer stays in the allocation list until finalize
    CUDA_TRY(cudaStreamSynchronize(e->stream));
  });
}
int lgarb_get_work_trace(lgarb_engine* e, int64_t col0, int64_t ncol, uint32_t* search_trips, uint32_t* other_kcycles) {
  LGARB_GUARD(e, {
    require_init(e);
    check_range(e, col0, ncol);
    if (!e->d_work_trace) throw std::runtime_error("work trace is off");
    d2h(e, search_trips, e->d_work_trace + col0, (size_t)ncol * sizeof(uint32_t));
    d2h(e, other_kcycles, e->d_work_trace + e->stride + col0, (size_t)ncol * sizeof(uint32_t));
  });
}
int lgarb_get_stage_profile(lgarb_engine* e, uint64_t* out15) {
  LGARB_GUARD(e, {
    require_init(e);
    if (!e->d_work_trace) throw std::runtime_error("work trace is off");
    d2h(e, out15, e->d_work_trace + 2 * e->stride, 15 * sizeof(uint64_t));
  });
}
int lgarb_get_warp_times(lgarb_engine* e, uint32_t* out16384) {
  LGARB_GUARD(e, {
    require_init(e);
    if (!e->d_work_trace) throw std::runtime_error("work trace is off");
    d2h(e, out16384, e->d_work_trace + 2 * e->stride + 64, 2 * 8192 * sizeof(uint32_t));
  });
}
int lgarb_set_stage_timing(lgarb_engine* e, int on) {
  LGARB_GUARD(e, {
    require_init(e);
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    if (e->side_stream) CUDA_TRY(cudaStreamSynchronize(e->side_stream));
    for (auto& se : e->stage_evs) {
      cudaEventDestroy(se.a);
      cudaEventDestroy(se.b);
    }
    e->stage_evs.clear();
    for (int s = 0; s <= LGAR_NSTAGE; s++) {
      e->stage_ms[s] = 0.0;
      e->stage_launches[s] = 0;
    }
    e->stage_timing = on != 0;
  });
}
int lgarb_get_stage_times(lgarb_engine* e, double* ms7, int64_t* launches7) {
  LGARB_GUARD(e, {
    require_init(e);
    if (!ms7 || !launches7) throw std::runtime_error("null result pointer");
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    if (e->side_stream) CUDA_TRY(cudaStreamSynchronize(e->side_stream));
    for (auto& se : e->stage_evs) {
      float ms = 0;
      CUDA_TRY(cudaEventElapsedTime(&ms, se.a, se.b));
      e->stage_ms[se.stage] += ms;
      e->stage_launches[se.stage] += 1;
      cudaEventDestroy(se.a);
      cudaEventDestroy(se.b);
    }
    e->stage_evs.clear();
    for (int s = 0; s <= LGAR_NSTAGE; s++) {
      ms7[s] = e->stage_ms[s];
      launches7[s] = e->stage_launches[s];
    }
  });
}
int lgarb_set_path_counters(lgarb_engine* e, int on) {
  LGARB_GUARD(e, {
    require_init(e);
    if (on && !e->d_paths) e->d_paths = e->dalloc<unsigned long long>(LGAR_NPATH);
    if (on) CUDA_TRY(cudaMemsetAsync(e->d_paths, 0, LGAR_NPATH * sizeof(unsigned long long), e->stream));
    if (!on) e->d_paths = nullptr;  // the buffer stays in the engine's allocation list until release
    CUDA_TRY(cudaStreamSynchronize(e->stream));
  });
}
int lgarb_get_path_counters(lgarb_engine* e, uint64_t* out, int n) {
  LGARB_GUARD(e, {
    require_init(e);
    if (!out || n < 0) throw std::runtime_error("bad arguments");
    if (!e->d_paths) throw std::runtime_error("path counters are off (lgarb_set_path_counters)");
    unsigned long long h[LGAR_NPATH];
    d2h(e, h, e->d_paths, sizeof(h));
    for (int i = 0; i < n; i++) out[i] = i < LGAR_NPATH ? (uint64_t)h[i] : 0;
  });
}
int lgarb_set_window_mode(lgarb_engine* e, int mode) {
  LGARB_GUARD(e, {
    // 2 (one kernel running the staged program per lane) and 4, 5 (whole steps per thread) were measured slower and removed
    if (mode != 0 && mode != 1 && mode != 3) throw std::runtime_error("window mode must be 0 (hour by hour), 1 (warp-tile window kernel) or 3 (stage kernels)");
    e->use_window = mode;
  });
}
int lgarb_set_wave_params(lgarb_engine* e, int search_quantum, int front_search_pairs, int finish_below) {
  LGARB_GUARD(e, {
    if (search_quantum > 0) e->wave_quantum = search_quantum;
    if (front_search_pairs > 0) e->wave_pairs = front_search_pairs;
    if (finish_below >= 0) e->wave_finish_below = finish_below;
  });
}
int64_t lgarb_get_launch_count(const lgarb_engine* e) { return e ? e->launches : -1; }
int64_t lgarb_get_last_active_count(lgarb_engine* e) {
  if (!e || !e->initialized) return -1;
  int32_t c = -1;
  try {
    d2h(e, &c, e->d.work_count, sizeof(int32_t));
  } catch (...) {
    return -1;
  }
  return c;
}
int lgarb_set_profiling(lgarb_engine* e, int on) {
  LGARB_GUARD(e, {
    require_init(e);
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    e->profiling = on != 0;
    e->ev_used = 0;
    e->prof_ms_stream = e->prof_ms_active = 0;
    e->prof_ms_window = 0;
    e->prof_windows = 0;
    e->prof_steps = 0;
    CUDA_TRY(cudaMemsetAsync(e->d_prof_acc, 0, 2 * sizeof(unsigned long long), e->stream));
  });
}

int lgarb_get_profile(lgarb_engine* e, double* out6) {
  LGARB_GUARD(e, {
    require_init(e);
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    for (size_t i = 0; i + 2 < e->ev_used; i += 3) {
      float a = 0, b = 0;
      CUDA_TRY(cudaEventElapsedTime(&a, e->ev_pool[i], e->ev_pool[i + 1]));
      CUDA_TRY(cudaEventElapsedTime(&b, e->ev_pool[i + 1], e->ev_pool[i + 2]));
      e->prof_ms_stream += a;
      e->prof_ms_active += b;
    }
    e->ev_used = 0;
    unsigned long long acc[2];
    d2h(e, acc, e->d_prof_acc, sizeof(acc));
    out6[0] = e->prof_ms_stream;
    out6[1] = e->prof_ms_active;
    out6[2] = (double)e->prof_steps;
    out6[3] = (double)acc[1];
    out6[4] = (double)acc[0];
    if (e->prof_windows > 0) {  // the last window's events have not been folded yet
      float ms = 0;
      CUDA_TRY(cudaEventElapsedTime(&ms, e->win_ev[0], e->win_ev[1]));
      e->prof_ms_window += ms;
      e->prof_windows = 0;
    }
    out6[5] = e->prof_ms_window;
  });
}

void* lgarb_get_stream(lgarb_engine* e) { return e ? (void*)e->stream : nullptr; }

int lgarb_debug_pow(lgarb_engine* e, int64_t n, const double* x, const double* y, double* out, int which) {
  LGARB_GUARD(e, {
    require_init(e);
    if (n < 0 || !x || !y || !out) throw std::runtime_error("bad arguments");
    if (n == 0) return LGARB_SUCCESS;
    double* dbuf = nullptr;
    CUDA_TRY(cudaMalloc((void**)&dbuf, (size_t)n * 3 * sizeof(double)));
    cudaError_t err = cudaMemcpyAsync(dbuf, x, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, e->stream);
    if (err == cudaSuccess) err = cudaMemcpyAsync(dbuf + n, y, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, e->stream);
    if (err == cudaSuccess) err = lgar_launch_debug_pow(dbuf, dbuf + n, dbuf + 2 * n, n, which, e->stream);
    if (err == cudaSuccess) err = cudaMemcpyAsync(out, dbuf + 2 * n, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, e->stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
    cudaFree(dbuf);
    e->launches += 1;
    CUDA_TRY(err);
  });
}

int lgarb_debug_functions(lgarb_engine* e, int fn, int layer, int64_t n, const double* a, const double* b, double* out) {
  LGARB_GUARD(e, {
    require_init(e);
    if (n < 0 || n > e->ncol || fn < 0 || fn > 11 || !a || !b || !out) throw std::runtime_error("bad arguments");
    if (n == 0) return LGARB_SUCCESS;
    double* dbuf = nullptr;
    CUDA_TRY(cudaMalloc((void**)&dbuf, (size_t)n * 3 * sizeof(double)));
    cudaError_t err = cudaMemcpyAsync(dbuf, a, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, e->stream);
    if (err == cudaSuccess) err = cudaMemcpyAsync(dbuf + n, b, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, e->stream);
    if (err == cudaSuccess) err = lgar_launch_debug_functions(e->cfg, e->d, fn, layer, n, dbuf, dbuf + n, dbuf + 2 * n, e->stream);
    if (err == cudaSuccess) err = cudaMemcpyAsync(out, dbuf + 2 * n, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, e->stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
    cudaFree(dbuf);
    e->launches += 1;
    CUDA_TRY(err);
  });
}

int lgarb_measure_fp64_peak(lgarb_engine* e, double ms, double* tflops) {
  LGARB_GUARD(e, {
    require_init(e);
    if (!tflops) throw std::runtime_error("null result pointer");
    int sms = 148;
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, e->device));
    const int blocks = sms * 8;
    double* sink = e->d_stage;  // any device double; the kernel never writes it in practice
    cudaEvent_t ev[2];
    CUDA_TRY(cudaEventCreate(&ev[0]));
    CUDA_TRY(cudaEventCreate(&ev[1]));
    auto run = [&](int iters) {
      CUDA_TRY(cudaEventRecord(ev[0], e->stream));
      CUDA_TRY(lgar_launch_dfma(sink, blocks, iters, e->stream));
      CUDA_TRY(cudaEventRecord(ev[1], e->stream));
      CUDA_TRY(cudaEventSynchronize(ev[1]));
      float t = 0;
      CUDA_TRY(cudaEventElapsedTime(&t, ev[0], ev[1]));
      return (double)t;
    };
    run(256);  // warm-up
    int iters = 2048;
    double t = run(iters);
    const double want = ms > 0 ? ms : 50.0;
    if (t > 0 && t < want) {
      iters = (int)std::min<double>(1e8, iters * want / t);
      t = run(iters);
    }
    cudaEventDestroy(ev[0]);
    cudaEventDestroy(ev[1]);
    e->launches += 3;
    *tflops = 2.0 * 64.0 * (double)iters * 256.0 * blocks / (t * 1e-3) / 1e12;
  });
}

}  // extern "C"
