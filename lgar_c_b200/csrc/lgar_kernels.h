// lgar_c_b200/csrc/lgar_kernels.h -- launch interface between the host engine and the step kernels.
#ifndef LGAR_KERNELS_H
#define LGAR_KERNELS_H

#ifndef LGAR_HOSTSIM
#include <cuda_runtime.h>
#endif

#include "lgar_types.h"

#define LGAR_ACTIVE_THREADS 128

// optional per-step output sinks: sinks.p[k] (device, may be null) receives output k of this step
// for every column, so a multi-step run can stream chosen variables into a [step][column] series.
struct LgarSinks {
  double* p[LGAR_NOUT];
};

// One launch of the window kernel: K consecutive forcing hours for every column.
struct LgarWindow {
  const double* precip;   // [K][series_stride] mm/h
  const double* pet;
  int64_t series_stride;
  LgarSinks sinks;        // each non-null: [K][sink_stride]
  int64_t sink_stride;
  int has_sinks;          // any sinks.p[k] non-null (host-computed: lets the per-hour code skip the output block)
  int32_t* nf_sink;       // nullable: [K][sink_stride] soil_num_wetting_fronts after each hour
  int K;
  int force_all_active;
  int32_t* tile_counter;  // device, zeroed before launch (tile counter / column counter)
  unsigned long long* prof;  // nullable: [0] += sum of front counts, [1] += active column-steps
  int slot_is_column;     // wavefront: slot i owns column i for the whole window (no column counter)
  // wavefront without slot_is_column: slots draw the entries 0 .. *col_count - 1 of col_list from the counter (null: the columns
  // 0 .. ncol - 1) -- the hour-by-hour Update() hands the work list of lgar_step_stream_kernel to the stage kernels this way
  const int32_t* col_list;
  const int32_t* col_count;
  // Forcing that is still on its way (lgarb_update_steps_host copies the series in chunks while the window runs): 0 = the whole
  // window is present, otherwise hours present so far + 1.  FETCH walks a column no further than that; a slot that gets there
  // waits in the FETCH queue (wave scheduler only).
  int frontier1;
  int32_t* waiting;       // nullable: += 1 for every slot FETCH left waiting at the frontier
  int32_t* done_columns;  // nullable: += 1 whenever a column finishes its window (wavefront scheduler)
  unsigned int* work_trace;  // nullable: [2][stride] per column: psi-search trips, kilo-cycles in the non-search stages (lanes kernel)
  int64_t trace_stride;
  unsigned long long* paths;  // nullable: [LGAR_NPATH] path counters (lgarb_set_path_counters)
  int use_ff;                 // psi searches inside a stage take the fast path of lgar_search.cuh (few slots in flight)
};
#define LGAR_TILE 256            // columns owned by one warp at a time
#define LGAR_WINDOW_THREADS 128
#define LGAR_MAX_WINDOW 32768    // hours per launch (time pointers are 16 bit)

// Wavefront scheduler state (mode 3): contexts of in-flight column-steps parked in HBM and one
// compacted queue pair per stage (see lgar_kernels.cu).
struct LgLane;
#define LGAR_NSTAGE 6
// stages of the resumable column program (lgar_lanes.cuh)
enum { LG_ST_FETCH = 0, LG_ST_HEAD = 1, LG_ST_FRONT = 2, LG_ST_SEARCH = 3, LG_ST_TAIL = 4, LG_ST_CORR = 5, LG_ST_DONE = 6 };
// Every queue is kept in LGAR_NBIN bins so that the lanes of a warp work on slots that look alike: HEAD / FRONT / TAIL slots are
// binned by their front count (LGAR_NBIN_N levels: the per-front loops of those stages then have the same trip count across the
// warp) and by flags (LGAR_NBIN_FLAG levels: HEAD -- it rains in the hour of the step; TAIL -- the step creates a surficial front,
// the list needs a correction; each selects a long branch), SEARCH slots by the layer of the front they search for (a trip
// evaluates theta in that many layers).
// A consumer walks the bins one after the other, each bin starting at a multiple of 32 of its index space (no warp straddles
// two bins).  Measured (profiles/r2_call8_summary.md): 1 bin 4.98e8, 2 bins 5.23e8, 4 bins 5.58e8, 6 x 2 bins 5.75e8 column-timesteps/s.
#ifndef LGAR_NBIN_N
#define LGAR_NBIN_N 6
#endif
#ifndef LGAR_NBIN_FLAG
#define LGAR_NBIN_FLAG 4
#endif
#define LGAR_NBIN (LGAR_NBIN_N * LGAR_NBIN_FLAG)
struct LgarWave {
  LgLane* ctx;                 // [nslots]
  int32_t* queue[LGAR_NSTAGE][2];  // [LGAR_NBIN][nslots] slot ids
  int32_t* qcount;             // [LGAR_NSTAGE][2][LGAR_NBIN]
  int32_t nslots;
  int32_t* done_columns;       // columns that finished their window
  int32_t* cursor;             // FETCH kernel: next queue entry to hand out (zero between launches)
  int32_t* waiting;            // slots FETCH left waiting for forcing since the last poll (the poll reads and clears it)
  int32_t* done_blocks;        // [LGAR_NSTAGE] blocks of the running consumer of a stage that have finished: the last one clears the consumed queue
};

#ifndef LGAR_HOSTSIM
// one launch of the kernel of `stage`, consuming queue[stage][consume]; append_mask bit s = buffer of stage s to append to;
// sm_count: multiprocessors of the device (grids are sized in multiples of it)
cudaError_t lgar_launch_wave_stage(const LgarConfig& cfg, const LgarDeviceState& d, const LgarWindow& w, const LgarWave& wv, int stage, int consume,
                                   unsigned append_mask, int quantum, int max_items, int sm_count, cudaStream_t stream);
// Finishing kernel with the lane records in shared memory (lgar_finish_smem.cu): every slot still in one of the stage queues is
// run to the end of the window.  slots_per_warp: 32 ... 2 records per warp; quantum: search trips per turn of the warp; use_ff:
// searches through the fast path of lgar_search.cuh.  The queue counts are cleared afterwards.
cudaError_t lgar_launch_wave_finish_smem(const LgarConfig& cfg, const LgarDeviceState& d, const LgarWindow& w, const LgarWave& wv, int in_flight,
                                         int slots_per_warp, int quantum, int use_ff, int sm_count, cudaStream_t stream);
cudaError_t lgar_launch_wave_finish(const LgarConfig& cfg, const LgarDeviceState& d, const LgarWindow& w, const LgarWave& wv, int in_flight,
                                    int lane_stride, int use_ff, int sm_count, cudaStream_t stream);
cudaError_t lgar_launch_wave_init(const LgarConfig& cfg, const LgarDeviceState& d, const LgarWave& wv, int slot_is_column, cudaStream_t stream);
// the 12 queue counts into pinned host memory (device-visible under unified addressing), no copy engine involved
cudaError_t lgar_launch_wave_poll(const LgarWave& wv, int32_t* host_counts, cudaStream_t stream);
size_t lgar_lane_context_bytes();
// mode 1: warp-tile window kernel
cudaError_t lgar_launch_window(const LgarConfig& cfg, const LgarDeviceState& d, const LgarWindow& w, int blocks, cudaStream_t stream, int mode);
int lgar_window_kernel_max_blocks_per_sm();

// the first kernel of that pair alone: cached-flux columns are finished, the others are listed in d.work / d.work_count[0]
cudaError_t lgar_launch_step_stream(const LgarConfig& cfg, const LgarDeviceState& d, const LgarSinks& sinks, int force_all_active, cudaStream_t stream);
// ev (nullable): three events recorded before the stream kernel, between the kernels, after the active kernel
cudaError_t lgar_launch_step(const LgarConfig& cfg, const LgarDeviceState& d, const LgarSinks& sinks, int active_blocks,
                             int force_all_active, cudaStream_t stream, cudaEvent_t* ev);
// sum of d.nfront over all columns, added into *acc (unsigned long long, device)
cudaError_t lgar_launch_sum_fronts(const LgarDeviceState& d, unsigned long long* acc, cudaStream_t stream);
cudaError_t lgar_launch_gather_fronts(const LgarDeviceState& d, int which, double scale, double* dst, int64_t col0, int64_t ncol,
                                      cudaStream_t stream);
cudaError_t lgar_launch_mass_balance(const LgarDeviceState& d, const double* volstart, const double* volchange, double* out, cudaStream_t stream);
cudaError_t lgar_launch_gather_layers(const LgarDeviceState& d, double* dst, int64_t col0, int64_t ncol, int L, cudaStream_t stream);
int lgar_active_kernel_max_blocks_per_sm();
// DFMA microbenchmark: `iters` trips of 8 independent FMA chains per thread on blocks x 256 threads; sink keeps the result alive
cudaError_t lgar_launch_dfma(double* sink, int blocks, int iters, cudaStream_t stream);
// out[i] = x[i]^y[i] with the physics' pow (which = 0: out-of-line function, 1: inlined common case + fall-back)
// leaf functions on a frozen snapshot: out[i] = fn(a[i], b[i]) for column i (see lgarb_debug_functions)
cudaError_t lgar_launch_debug_functions(const LgarConfig& cfg, const LgarDeviceState& d, int fn, int layer, int64_t n, const double* a, const double* b,
                                        double* out, cudaStream_t stream);
cudaError_t lgar_launch_debug_pow(const double* x, const double* y, double* out, int64_t n, int which, cudaStream_t stream);
#endif

#endif
