// lgar_c_b200/csrc/lgar_kernels.cu -- the batched BmiLGAR::Update() step for sm_100a.
//
// One forcing step over all columns is two launches:
//
//   lgar_step_stream_kernel   every column, one thread each, fully coalesced SoA traffic.
//                             Runs the adaptive-dt / flux-cache decision (reference
//                             src/bmi_lgar.cxx:260-318).  Columns that replay cached fluxes (80-90 %
//                             of column-steps on the shipped configs) and
//                             invalid-soil columns are finished here without touching their fronts:
//                             reservoir, GIUH, mass balance, outputs.  Columns that must compute
//                             fluxes are compacted into a work list (warp-aggregated atomics).
//   lgar_step_active_kernel   work-list entries only, one thread per active column, so warps are
//                             dense in the FP64-heavy path instead of 15 % occupied.  Warps pull
//                             32 entries at a time from a device-side cursor (heavy-tailed trip
//                             counts: a static split would leave SMs idle).  Fronts live in
//                             thread-private arrays for the duration of the step.
//
// HBM-bound in the stream kernel, FP64-pipe-bound in the active kernel; no tensor-core work exists
// on this path (nothing is a contraction).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstddef>
#include <cstdio>
#include <cstdlib>

#include "lgar_lanes.cuh"


__global__ void __launch_bounds__(256) lgar_step_stream_kernel(const __grid_constant__ LgarConfig cfg, const __grid_constant__ LgarDeviceState d, const __grid_constant__ LgarSinks sinks, int force_all_active) {
  const int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool needs_active = false;
  if (col < d.ncol) needs_active = lg_stream_column(cfg, d, sinks, col, force_all_active);
  // warp-aggregated append to the work list
  const unsigned mask = __ballot_sync(0xFFFFFFFFu, needs_active);
  if (mask) {
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(d.work_count, __popc(mask));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    if (needs_active) d.work[base + __popc(mask & ((1u << lane) - 1))] = (int32_t)col;
  }
}

__global__ void __launch_bounds__(LGAR_ACTIVE_THREADS) lgar_step_active_kernel(const __grid_constant__ LgarConfig cfg, const __grid_constant__ LgarDeviceState d, const __grid_constant__ LgarSinks sinks) {
  const int lane = threadIdx.x & 31;
  const int count = d.work_count[0];
  for (;;) {
    int base = 0;
    if (lane == 0) base = atomicAdd(d.work_count + 1, 32);
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (base >= count) break;
    const int idx = base + lane;
    if (idx < count) lg_active_column(cfg, d, sinks, d.work[idx]);
  }
}

// ---- window kernel ----------------------------------------------------------------------------------
// K forcing hours per launch, no grid-wide or block-wide barrier anywhere.  A warp owns a tile of
// LGAR_TILE consecutive columns at a time (tiles are handed out by an atomic counter, so SMs stay
// busy until the last tile).  Every column of the tile carries its own time pointer:
//   ADVANCE  each lane walks its columns through the hours that need no flux computation (scalars in
//            registers across hours, fronts untouched) until the column needs an active step;
//   COLLECT  the columns now waiting at an active step are compacted into a ready list (ballot);
//   PROCESS  the ready list is worked off 32 columns at a time, one full Update() per lane.
// After the first round every column of the tile sits at an active step of its own hour, so PROCESS
// batches are full warps.  A straggler (the trip counts are heavy tailed: median 7, max > 20 000
// theta evaluations per active step) delays only its own tile, never the other 147 SMs.
__global__ void __launch_bounds__(LGAR_WINDOW_THREADS) lgar_window_kernel(const __grid_constant__ LgarConfig cfg, const __grid_constant__ LgarDeviceState d, const __grid_constant__ LgarWindow w) {
  __shared__ unsigned short s_t[LGAR_WINDOW_THREADS / 32][LGAR_TILE];
  __shared__ unsigned short s_ready[LGAR_WINDOW_THREADS / 32][LGAR_TILE];
  __shared__ unsigned char s_flag[LGAR_WINDOW_THREADS / 32][LGAR_TILE];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  unsigned short* tcur = s_t[wid];
  unsigned short* ready = s_ready[wid];
  unsigned char* flag = s_flag[wid];   // 0 idle, 1 ready, 2 done
  const int64_t ntiles = (d.ncol + LGAR_TILE - 1) / LGAR_TILE;
  unsigned long long front_sum = 0, active_sum = 0;
  unsigned long long* fsp = w.prof ? &front_sum : nullptr;
  for (;;) {
    int tile = 0;
    if (lane == 0) tile = atomicAdd(w.tile_counter, 1);
    tile = __shfl_sync(0xFFFFFFFFu, tile, 0);
    if (tile >= ntiles) break;
    const int64_t c0 = (int64_t)tile * LGAR_TILE;
    for (int j = lane; j < LGAR_TILE; j += 32) {
      tcur[j] = 0;
      flag[j] = (c0 + j < d.ncol) ? 0 : 2;
    }
    __syncwarp();
    for (;;) {
      for (int j = lane; j < LGAR_TILE; j += 32) {  // ADVANCE
        if (flag[j] == 0) {
          const int t = lg_advance_column(cfg, d, w, c0 + j, tcur[j], fsp);
          tcur[j] = (unsigned short)t;
          flag[j] = (t < w.K) ? 1 : 2;
        }
      }
      __syncwarp();
      int n = 0;
      for (int base = 0; base < LGAR_TILE; base += 32) {  // COLLECT
        const int j = base + lane;
        const bool r = flag[j] == 1;
        const unsigned m = __ballot_sync(0xFFFFFFFFu, r);
        if (r) ready[n + __popc(m & ((1u << lane) - 1))] = (unsigned short)j;
        n += __popc(m);
      }
      __syncwarp();
      if (n == 0) break;
      active_sum += (lane == 0) ? n : 0;
      for (int b = 0; b < n; b += 32) {  // PROCESS
        const int i = b + lane;
        if (i < n) {
          const int j = ready[i];
          lg_active_column_at(cfg, d, w, c0 + j, tcur[j], fsp);
          tcur[j] = (unsigned short)(tcur[j] + 1);
          flag[j] = 0;
        }
        __syncwarp();
      }
    }
  }
  if (w.prof) {
    for (int o = 16; o > 0; o >>= 1) front_sum += __shfl_down_sync(0xFFFFFFFFu, front_sum, o);
    if (lane == 0) {
      atomicAdd(w.prof, front_sum);
      atomicAdd(w.prof + 1, active_sum);
    }
  }
}

// ---- wavefront kernels (mode 3) -------------------------------------------------------------------------
// A single kernel running the staged program of lgar_lanes.cuh per lane (round 1's "lanes kernel", removed in round 2) proved
// the program right but ran at 0.2 instructions/clock/scheduler: its code (206-437 KB of SASS) does not fit the 32 KB L1.5
// instruction cache and every warp sits somewhere else in it (ncu: 26 of 34 stall cycles per instruction are instruction
// fetch).  Here every stage is
// its own small kernel; the context of an in-flight column-step (LgLane) is parked in HBM between
// stages and stages hand slots to each other through compacted queues, so every kernel runs full warps
// over one short piece of code.  The host enqueues rounds of FETCH, HEAD, (FRONT, SEARCH) x n, TAIL
// without looking at the queues and polls one counter now and then.  A queue pair per stage: a launch
// of stage S consumes the buffer that has been appended to since S last ran, and everything produced
// meanwhile (including S re-queuing itself) goes to the other one.
#include "lgar_ctx.cuh"

// Resident blocks per SM the compiler must leave room for, per stage (0 = no constraint).  The stages are bound by
// memory latency, so occupancy pays more than a few extra registers; values picked from the sweep in
// profiles/r1_ncu_summary.md.
#ifndef LGAR_MB_FETCH
#define LGAR_MB_FETCH 0
#endif
#ifndef LGAR_MB_HEAD
#define LGAR_MB_HEAD 0
#endif
#ifndef LGAR_MB_FRONT
#define LGAR_MB_FRONT 0
#endif
#ifndef LGAR_MB_SEARCH
#define LGAR_MB_SEARCH 0
#endif
#ifndef LGAR_MB_TAIL
#define LGAR_MB_TAIL 0
#endif
#ifndef LGAR_MB_CORR
#define LGAR_MB_CORR 0
#endif
constexpr int wave_min_blocks(int stage) {
  return stage == LG_ST_FETCH ? LGAR_MB_FETCH : stage == LG_ST_HEAD ? LGAR_MB_HEAD : stage == LG_ST_FRONT ? LGAR_MB_FRONT
       : stage == LG_ST_SEARCH ? LGAR_MB_SEARCH : stage == LG_ST_TAIL ? LGAR_MB_TAIL : LGAR_MB_CORR;
}
// FF: the SEARCH / CORR stages take the fast path of lgar_search.cuh (few slots in flight: a latency tool).  A template parameter,
// not an argument: the plain instantiation must not carry the fast path's registers (occupancy of the bulk rounds).
template <int STAGE, bool FF = false>
__global__ void __launch_bounds__(128, wave_min_blocks(STAGE)) lgar_wave_kernel(const __grid_constant__ LgarConfig cfg, const __grid_constant__ LgarDeviceState d, const __grid_constant__ LgarWindow w, const __grid_constant__ LgarWave wv, int consume,
                                                        unsigned append_mask, int quantum) {
  __shared__ WaveItems items;
  wave_items_init(items, wv, STAGE, consume);
  LgLaneSlot* slots = reinterpret_cast<LgLaneSlot*>(wv.ctx);
  unsigned long long front_sum = 0, active_sum = 0;
  unsigned long long* fsp = w.prof ? &front_sum : nullptr;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < items.voff[LGAR_NBIN]; i += gridDim.x * blockDim.x) {
    const int slot_or_pad = wave_item(wv, items, STAGE, consume, i);
    const bool valid = slot_or_pad >= 0;
    const int slot = valid ? slot_or_pad : 0;
    int next = -1, bin = 0;  // bin of the queue the slot goes to next: by front count; SEARCH: by the layer of the searched front
    if (STAGE == LG_ST_SEARCH) {
      if (valid) {  // only the search record travels
        LgLane& G = slots[slot].L;
        LgSearch q = G.q;
        const LgarColumnClass* cls = G.cls;
        lg_search_some(q, *cls, quantum, FF);
        G.q = q;
        if (q.done) G.resume = true;
        next = q.done ? LG_ST_FRONT : LG_ST_SEARCH;
        bin = q.done ? wave_bin_fronts(G.nf_stored) : wave_bin_layers(q.layer_num);
      }
    } else if (STAGE == LG_ST_CORR) {
      if (valid) {  // the correction record and the front list travel
        LgLane& G = slots[slot].L;
        LgCorr k = G.k;
        // the correction search reads depth/theta/layer/to_bottom of the live fronts and rewrites theta and psi
        struct alignas(32) FrontsBuf { LgFronts f; } fb;
        char* l = reinterpret_cast<char*>(&fb);
        char* g = reinterpret_cast<char*>(&G.f);
        ctx_load<kFrontTailBytes>(reinterpret_cast<FrontsBuf*>(l + kFrontTail), reinterpret_cast<const FrontsBuf*>(g + kFrontTail));
        const int cf = (fb.f.n + 3) >> 2;
#pragma unroll
        for (int a = 0; a < 3; a++) ctx_load_n(l + kPitch * a, g + kPitch * a, cf);
        const LgarColumnClass* cls = G.cls;
        lg_corr_some(k, fb.f, *cls, quantum, FF);
        G.k = k;
#pragma unroll
        for (int a = 1; a < 3; a++) ctx_store_n(g + kPitch * a, l + kPitch * a, cf);
        next = k.done ? LG_ST_TAIL : LG_ST_CORR;
        bin = k.done ? wave_bin_fronts(fb.f.n, wave_tail_flags(G)) : 0;
      }
    } else if (STAGE == LG_ST_FETCH) {
      if (valid) {  // works on the record in place: scalars and epilogue state through registers, nothing else is touched
        LgLane& G = slots[slot].L;
        lg_stage_fetch(cfg, d, w, G, fsp);
        next = G.stage;
        bin = wave_bin_fronts(G.f.n, (next == LG_ST_HEAD && w.precip[(int64_t)G.t * w.series_stride + G.col] > 0.0) ? 1 : 0);
      }
    } else if (valid) {
      LgLaneSlot S;
      LgLane& L = S.L;
      // What each stage moves (parts of the record: lgar_ctx.cuh).  HEAD: status word, scalars, fronts in (the epilogue state stays
      // where it is unless the test mode that can finish a step here is on); everything it has built out.  FRONT: status word,
      // fronts, the 128 bytes of step context the front loop works on, state_previous (in only), the search record.  TAIL:
      // the whole state and step context in; the state out when the step is over, state + step context when it goes on.
      constexpr int kHeadIn = LG_PART_CTX | LG_PART_S | LG_PART_FRONTS;
      constexpr int kFrontIO = LG_PART_CTX | LG_PART_FRONTS | LG_PART_STEPF;
      constexpr int kGoOn = LG_PART_CTX | LG_PART_S | LG_PART_FRONTS | LG_PART_STEP;
      if (STAGE == LG_ST_HEAD) {  // HEAD starts the step context behind the resident column state
        if (w.force_all_active)
          lane_load<LG_PART_STATE>(&S, &slots[slot]);
        else
          lane_load<kHeadIn>(&S, &slots[slot]);
        lg_stage_head(cfg, d, w, L, fsp);
        active_sum += (L.stage == LG_ST_FRONT) ? 1 : 0;
        // every new step goes straight into its front loop: saves one context round trip per step
        if (L.stage == LG_ST_FRONT) lg_stage_front(L);
      } else {
        if (STAGE == LG_ST_TAIL)
          lane_load<LG_PART_STATE | LG_PART_STEP>(&S, &slots[slot]);
        else
          lane_load<kFrontIO | LG_PART_OLD | LG_PART_SEARCH>(&S, &slots[slot]);
        L.c.cfg = &cfg;
        L.c.cls = L.cls;
        L.c.f = &L.f;
        L.c.ff = L.ffv;
        L.c.paths = w.paths;
        L.c.use_ff = w.use_ff != 0;
        if (STAGE == LG_ST_FRONT)
          lg_stage_front(L);
        else
          lg_stage_tail(cfg, d, w, L, fsp);
      }
      next = L.stage;
      bin = next == LG_ST_SEARCH ? wave_bin_layers(L.q.layer_num) : (next == LG_ST_CORR || next == LG_ST_FETCH) ? 0 : wave_bin_fronts(L.f.n, next == LG_ST_TAIL ? wave_tail_flags(L) : 0);
      if (STAGE == LG_ST_FRONT) {
        if (next == LG_ST_SEARCH)
          lane_store<kFrontIO | LG_PART_SEARCH>(&slots[slot], &S);
        else
          lane_store<kFrontIO>(&slots[slot], &S);  // the front loop is over
      } else if (next == LG_ST_FETCH) {
        lane_store<LG_PART_STATE>(&slots[slot], &S);  // the step is over: only the column state lives on
      } else if (next == LG_ST_TAIL || next == LG_ST_CORR) {
        lane_store<kGoOn>(&slots[slot], &S);  // HEAD -> TAIL (no search), TAIL -> CORR -> TAIL: neither state_previous nor the psi search is live
      } else if (next == LG_ST_FRONT) {
        lane_store<kGoOn | LG_PART_OLD>(&slots[slot], &S);  // TAIL has begun the next sub-cycle
      } else {
        lane_store<kGoOn | LG_PART_OLD | LG_PART_SEARCH>(&slots[slot], &S);  // HEAD -> SEARCH
      }
    }
    wave_push(wv, append_mask, valid ? next : -1, bin, slot);
  }
  if (w.prof && (STAGE == LG_ST_FETCH || STAGE == LG_ST_HEAD || STAGE == LG_ST_TAIL)) {
    if (front_sum) atomicAdd(w.prof, front_sum);
    if (active_sum) atomicAdd(w.prof + 1, active_sum);
  }
  wave_consumed(wv, STAGE, consume, false);
}

// ---- FETCH with per-lane refill -------------------------------------------------------------------------
// FETCH walks a column through the hours that replay cached fluxes until it needs an active step.  The walks are heavy
// tailed (a dry spell is tens to hundreds of hours, a wet one a single hour), so with one queue entry per thread a warp
// waits for its longest walk: ncu 3.8 of 32 lanes per instruction.  Here a lane that has finished its walk takes the next
// queue entry while its neighbours keep walking (one warp-aggregated atomicAdd on a cursor per refill), so a warp's time is
// the MEAN walk of what it handles, not the maximum; the forcing of the next hour is requested before the current hour is
// computed.  The arithmetic per hour is lg_resident_hour(), the body lg_advance_resident() loops over: bit-identical.
__global__ void __launch_bounds__(128) lgar_wave_fetch_kernel(const __grid_constant__ LgarConfig cfg, const __grid_constant__ LgarDeviceState d, const __grid_constant__ LgarWindow w,
                                                              const __grid_constant__ LgarWave wv, int consume, unsigned append_mask) {
  const unsigned FULL = 0xFFFFFFFFu;
  const int n = wv.qcount[(LG_ST_FETCH * 2 + consume) * LGAR_NBIN];  // everything that goes to FETCH goes to its bin 0
  const int lane = threadIdx.x & 31;
  const int K = w.K, Kp = lg_hours_present(w);  // end of the window, end of the forcing that has arrived
  LgLaneSlot* slots = reinterpret_cast<LgLaneSlot*>(wv.ctx);
  unsigned long long front_sum = 0;
  unsigned long long* fsp = w.prof ? &front_sum : nullptr;
  bool have = false, exhausted = false;
  int slot = 0, t = 0, nf = 0;
  int64_t col = 0;
  LgScalars s;
  LgEpi ep;
  double Pn = 0.0, En = 0.0;  // forcing of hour t, requested one hour ahead
  bool rain = false;          // it rains in the hour the slot stopped at (bin of the HEAD queue)
  for (;;) {
    if (!exhausted) {
      const unsigned need = __ballot_sync(FULL, !have);
      if (need) {
        int base = 0;
        if (lane == 0) base = atomicAdd(wv.cursor, __popc(need));
        base = __shfl_sync(FULL, base, 0);
        exhausted = base + __popc(need) >= n;
        const int idx = base + __popc(need & ((1u << lane) - 1));
        if (!have && idx < n) {
          slot = wv.queue[LG_ST_FETCH][consume][idx];
          LgLane& G = slots[slot].L;
          col = G.col;
          t = G.t;
          nf = G.f.n;
          s = G.s;
          ep = G.ep;
          have = true;
          if (t < Kp && G.cls->invalid_soil_type) t = lg_advance_resident(cfg, d, w, G.cls, col, t, nf, s, ep, fsp);  // Q = precip: no walk
          if (t < Kp) {
            Pn = w.precip[(int64_t)t * w.series_stride + col];
            En = w.pet[(int64_t)t * w.series_stride + col];
          }
        }
      }
    }
    if (!__any_sync(FULL, have)) break;
    int next = -1;
    if (have) {
      bool ended = t >= Kp;
      if (!ended) {
        const double P = Pn, E = En;
        rain = P > 0.0;
        if (t + 1 < Kp) {
          Pn = w.precip[(int64_t)(t + 1) * w.series_stride + col];
          En = w.pet[(int64_t)(t + 1) * w.series_stride + col];
        }
        if (lg_resident_hour(cfg, d, w, col, t, nf, s, ep, fsp, P, E)) {
          t++;
          ended = t >= Kp;
        } else {
          ended = true;
        }
      }
      if (ended) {
        LgLane& G = slots[slot].L;
        G.s = s;
        G.ep = ep;
        G.t = t;
        if (t < Kp) {
          G.stage = LG_ST_HEAD;
          next = LG_ST_HEAD;
        } else if (t < K) {  // the forcing of hour t has not arrived yet: wait in the FETCH queue
          if (w.waiting) atomicAdd(w.waiting, 1);
          G.stage = LG_ST_FETCH;
          next = LG_ST_FETCH;
        } else {  // this column has finished the window (lg_stage_fetch, slot == column)
          lg_lane_export(cfg, d, col, G);
          if (w.done_columns) atomicAdd(w.done_columns, 1);
          if (w.work_trace) {
            w.work_trace[col] = G.tr_iters;
            w.work_trace[w.trace_stride + col] = G.tr_kcyc;
          }
          G.col = -1;
          G.stage = LG_ST_DONE;
          next = LG_ST_DONE;
        }
        have = false;
      }
    }
    wave_push(wv, append_mask, (next == LG_ST_HEAD || next == LG_ST_FETCH) ? next : -1, next == LG_ST_HEAD ? wave_bin_fronts(nf, rain ? 1 : 0) : 0, slot);
  }
  if (w.prof && front_sum) atomicAdd(w.prof, front_sum);
  wave_consumed(wv, LG_ST_FETCH, consume, true);
}

// Finishing kernel: when only a few slots are left in flight (the columns that are active in almost every
// hour of the window), a round of ~13 launches costs ~2.5 ms whatever it processes.  Here every remaining
// slot is taken by one thread which runs its column through all stages to the end of the window without
// going back to HBM between stages: the critical path per active step drops from one round to the serial
// execution time of the step.  Lane utilisation is poor by construction, which does not matter for the few
// thousand slots this is used for.
#ifndef LGAR_MB_FINISH
#define LGAR_MB_FINISH 0
#endif
__global__ void __launch_bounds__(128, LGAR_MB_FINISH) lgar_wave_finish_kernel(const __grid_constant__ LgarConfig cfg, const __grid_constant__ LgarDeviceState d, const __grid_constant__ LgarWindow w, const __grid_constant__ LgarWave wv, int lane_stride, int use_ff) {
  LgLaneSlot* slots = reinterpret_cast<LgLaneSlot*>(wv.ctx);
  int total = 0;
  for (int q = 0; q < 2 * LGAR_NSTAGE * LGAR_NBIN; q++) total += wv.qcount[q];
  unsigned long long front_sum = 0, active_sum = 0;
  unsigned long long* fsp = w.prof ? &front_sum : nullptr;
  // `lane_stride` > 1: only every lane_stride-th lane of a warp takes a slot.  The lanes of a warp sit in different
  // stages of different columns and are executed one after the other, so the time of the kernel -- the serial
  // chain of the column with the most active hours left -- shrinks with the number of slots per warp as long as
  // the warps still fit the machine.
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gtid % lane_stride != 0) return;
  for (int i = gtid / lane_stride; i < total; i += gridDim.x * blockDim.x / lane_stride) {
    int stage = LG_ST_DONE;
    const int slot = wave_item_any(wv, i, &stage);  // item i of the concatenation of all queues
    if (slot < 0) continue;
    LgLaneSlot S;
    lane_load<LG_PART_ALL>(&S, &slots[slot]);
    LgLane& L = S.L;
    L.stage = stage;  // the queue a slot sits in is authoritative: the partial-context stages do not rewrite L.stage
    L.c.cfg = &cfg;
    L.c.cls = L.cls;
    L.c.f = &L.f;
    L.c.ff = L.ffv;
    L.c.paths = w.paths;
    L.c.use_ff = use_ff != 0;  // latency regime: see lg_search_some
    while (L.stage != LG_ST_DONE) {
      switch (L.stage) {
        case LG_ST_FETCH: lg_stage_fetch(cfg, d, w, L, fsp); break;
        case LG_ST_HEAD:
          lg_stage_head(cfg, d, w, L, fsp);
          active_sum += (L.stage == LG_ST_FRONT) ? 1 : 0;
          break;
        case LG_ST_FRONT: lg_stage_front(L); break;
        case LG_ST_SEARCH:
          lg_search_run(L.q, *L.cls, use_ff != 0);
          L.resume = true;
          L.stage = LG_ST_FRONT;
          break;
        case LG_ST_TAIL: lg_stage_tail(cfg, d, w, L, fsp); break;
        case LG_ST_CORR:
          lg_corr_run(L.k, L.f, *L.cls, use_ff != 0);
          L.stage = LG_ST_TAIL;
          break;
        default: L.stage = LG_ST_DONE; break;
      }
    }
  }
  if (w.prof) {
    if (front_sum) atomicAdd(w.prof, front_sum);
    if (active_sum) atomicAdd(w.prof + 1, active_sum);
  }
}

// Queue counts to the host through mapped pinned memory: a cudaMemcpyAsync of these 48 bytes would queue behind whatever
// bulk device-to-host copy (output series of the previous chunk of a streamed call) occupies the copy engine.
__global__ void lgar_wave_poll_kernel(const int32_t* qcount, int32_t* waiting, int32_t* host_counts) {
  if (threadIdx.x < 2 * LGAR_NSTAGE) {  // the host sees one count per queue: the sum over its bins
    int n = 0;
    for (int b = 0; b < LGAR_NBIN; b++) n += qcount[threadIdx.x * LGAR_NBIN + b];
    host_counts[threadIdx.x] = n;
  }
  if (threadIdx.x == 2 * LGAR_NSTAGE) {  // slots left waiting for forcing since the last poll
    host_counts[threadIdx.x] = *waiting;
    *waiting = 0;
  }
  __threadfence_system();
}
cudaError_t lgar_launch_wave_poll(const LgarWave& wv, int32_t* host_counts, cudaStream_t stream) {
  lgar_wave_poll_kernel<<<1, 32, 0, stream>>>(wv.qcount, wv.waiting, host_counts);
  return cudaGetLastError();
}

__global__ void lgar_wave_clear_kernel(LgarWave wv) {
  for (int i = threadIdx.x; i < 2 * LGAR_NSTAGE * LGAR_NBIN; i += blockDim.x) wv.qcount[i] = 0;
}

cudaError_t lgar_launch_wave_finish(const LgarConfig& cfg, const LgarDeviceState& d, const LgarWindow& w, const LgarWave& wv, int in_flight,
                                    int lane_stride, int use_ff, int sm_count, cudaStream_t stream) {
  lane_stride = std::max(1, std::min(lane_stride, 32));
  const int blocks = std::max(1, std::min((int)(((int64_t)in_flight * lane_stride + 127) / 128), sm_count * 16));
  lgar_wave_finish_kernel<<<blocks, 128, 0, stream>>>(cfg, d, w, wv, lane_stride, use_ff);
  lgar_wave_clear_kernel<<<1, 256, 0, stream>>>(wv);
  return cudaGetLastError();
}

__global__ void lgar_wave_init_kernel(const __grid_constant__ LgarConfig cfg, const __grid_constant__ LgarDeviceState d, const __grid_constant__ LgarWave wv,
                                      int slot_is_column) {
  LgLaneSlot* slots = reinterpret_cast<LgLaneSlot*>(wv.ctx);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < wv.nslots; i += gridDim.x * blockDim.x) {
    LgLane& L = slots[i].L;
    L.stage = LG_ST_FETCH;
    // slot == column: the slot imports its column's state here (coalesced reads of the SoA arrays) and keeps it
    // for the whole window; with fewer slots than columns, slots draw columns from a counter in FETCH instead
    L.col = slot_is_column ? i : -1;
    L.t = 0;
    L.nf_stored = 0;
    L.nold_stored = 0;
    L.resume = false;
    L.tr_iters = 0;
    L.tr_kcyc = 0;
    if (slot_is_column) lg_lane_import(cfg, d, i, L);
    wv.queue[LG_ST_FETCH][0][i] = i;
  }
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < LGAR_NSTAGE * 2 * LGAR_NBIN; i += blockDim.x) wv.qcount[i] = (i == LG_ST_FETCH * 2 * LGAR_NBIN) ? wv.nslots : 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    *wv.done_columns = 0;
    *wv.waiting = 0;
  }
}

cudaError_t lgar_launch_wave_init(const LgarConfig& cfg, const LgarDeviceState& d, const LgarWave& wv, int slot_is_column, cudaStream_t stream) {
  lgar_wave_init_kernel<<<(wv.nslots + 255) / 256, 256, 0, stream>>>(cfg, d, wv, slot_is_column);
  return cudaGetLastError();
}

cudaError_t lgar_launch_wave_stage(const LgarConfig& cfg, const LgarDeviceState& d, const LgarWindow& w, const LgarWave& wv, int stage, int consume,
                                   unsigned append_mask, int quantum, int max_items, int sm_count, cudaStream_t stream) {
  // grid-stride over the queue; max_items (slots in flight at the start of the round) bounds every queue
  const int blocks = std::max(1, std::min((std::min(wv.nslots, max_items) + 127) / 128, sm_count * 16));
  const size_t sm = 0;
  static const int fetch_refill = getenv("LGARB_FETCH_REFILL") ? atoi(getenv("LGARB_FETCH_REFILL")) : 8;  // 0: one queue entry per thread; n: refill kernel, n resident blocks per SM
  const bool refill = stage == LG_ST_FETCH && fetch_refill && w.slot_is_column;
  switch (stage) {
    case LG_ST_FETCH:
      if (refill)
        lgar_wave_fetch_kernel<<<std::min(blocks, sm_count * fetch_refill), 128, 0, stream>>>(cfg, d, w, wv, consume, append_mask);
      else
        lgar_wave_kernel<LG_ST_FETCH><<<blocks, 128, sm, stream>>>(cfg, d, w, wv, consume, append_mask, quantum);
      break;
    case LG_ST_HEAD: lgar_wave_kernel<LG_ST_HEAD><<<blocks, 128, sm, stream>>>(cfg, d, w, wv, consume, append_mask, quantum); break;
    case LG_ST_FRONT: lgar_wave_kernel<LG_ST_FRONT><<<blocks, 128, sm, stream>>>(cfg, d, w, wv, consume, append_mask, quantum); break;
    case LG_ST_SEARCH:
      if (w.use_ff) lgar_wave_kernel<LG_ST_SEARCH, true><<<blocks, 128, sm, stream>>>(cfg, d, w, wv, consume, append_mask, quantum);
      else lgar_wave_kernel<LG_ST_SEARCH><<<blocks, 128, sm, stream>>>(cfg, d, w, wv, consume, append_mask, quantum);
      break;
    case LG_ST_CORR:
      if (w.use_ff) lgar_wave_kernel<LG_ST_CORR, true><<<blocks, 128, sm, stream>>>(cfg, d, w, wv, consume, append_mask, quantum);
      else lgar_wave_kernel<LG_ST_CORR><<<blocks, 128, sm, stream>>>(cfg, d, w, wv, consume, append_mask, quantum);
      break;
    default: lgar_wave_kernel<LG_ST_TAIL><<<blocks, 128, sm, stream>>>(cfg, d, w, wv, consume, append_mask, quantum); break;
  }
  return cudaGetLastError();  // the consumed queue is cleared by the last block of the kernel itself (wave_consumed)
}

size_t lgar_lane_context_bytes() { return sizeof(LgLaneSlot); }

// FP64 pipe microbenchmark: 8 independent fused multiply-add chains per thread (asm: -fmad=false must not split them)
__global__ void __launch_bounds__(256) lgar_dfma_kernel(double* sink, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 0.999999, c = 1e-7;
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      asm volatile("fma.rn.f64 %0, %0, %8, %9;\n\tfma.rn.f64 %1, %1, %8, %9;\n\tfma.rn.f64 %2, %2, %8, %9;\n\tfma.rn.f64 %3, %3, %8, %9;\n\t"
                   "fma.rn.f64 %4, %4, %8, %9;\n\tfma.rn.f64 %5, %5, %8, %9;\n\tfma.rn.f64 %6, %6, %8, %9;\n\tfma.rn.f64 %7, %7, %8, %9;"
                   : "+d"(a0), "+d"(a1), "+d"(a2), "+d"(a3), "+d"(a4), "+d"(a5), "+d"(a6), "+d"(a7)
                   : "d"(m), "d"(c));
    }
  }
  const double r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (r == 123.456) sink[0] = r;
}
cudaError_t lgar_launch_dfma(double* sink, int blocks, int iters, cudaStream_t stream) {
  lgar_dfma_kernel<<<blocks, 256, 0, stream>>>(sink, iters);
  return cudaGetLastError();
}

// x^y as the physics evaluates it, element by element (lgarb_debug_pow: compared with the host's pow by the tests)
__global__ void lgar_debug_pow_kernel(const double* __restrict__ x, const double* __restrict__ y, double* __restrict__ out, int64_t n, int which) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = which ? lg_pow_hot(x[i], y[i]) : lg_pow(x[i], y[i]);
}
cudaError_t lgar_launch_debug_pow(const double* x, const double* y, double* out, int64_t n, int which, cudaStream_t stream) {
  lgar_debug_pow_kernel<<<1184, 256, 0, stream>>>(x, y, out, n, which);
  return cudaGetLastError();
}

// Frozen-snapshot evaluation of the leaf functions of the path (lgarb_debug_functions): one item per
// column, the column's class row and -- for the list functions -- its current fronts and reservoir storages.
__global__ void lgar_debug_functions_kernel(const __grid_constant__ LgarConfig cfg, const __grid_constant__ LgarDeviceState d, int fn, int layer, int64_t n,
                                            const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const LgarColumnClass& cls = d.classes[d.class_id[i]];
    const int l = layer < 1 ? 1 : (layer > cls.num_layers ? cls.num_layers : layer);
    const LgarLayer& s = cls.lay[l];
    double r = LG_NAN;
    if (fn == 0) r = lg_theta_from_h(a[i], s);                    // calc_theta_from_h, soil_funcs.cxx:147-150
    else if (fn == 1) r = lg_Se_from_h(a[i], s);                  // calc_Se_from_h, :155-159
    else if (fn == 2) r = lg_K_from_Se(a[i], s, s.Ksat);          // calc_K_from_Se, :164-167
    else if (fn == 3) r = lg_h_from_Se(a[i], s);                  // calc_h_from_Se, :172-179
    else if (fn == 4) r = lg_Geff(cfg, a[i], b[i], s, s.Ksat);    // calc_Geff, :15-131
    else if (fn == 5 || fn == 6) {
      LgFronts f;
      load_fronts(d, i, f);
      LgCtx c;
      c.cfg = &cfg;
      c.cls = &cls;
      c.f = &f;
      c.status = 0;
      c.ff = lg_ff_ones;
      c.paths = nullptr;
      c.use_ff = false;
      r = (fn == 5) ? lg_mass_bal(c) : lg_calc_aet(c, a[i], b[i]);  // lgar_calc_mass_bal lgar.cxx:2632-2668; calc_aet aet.cxx:17-77
    } else if (fn >= 7 && fn <= 11) {                              // calc_CR_Q, conceptual_reservoir.cxx:7-45
      LgScalars sc;
      load_scalars(d, i, sc);
      double fast = sc.cr_fast, slow = sc.cr_slow;
      if (fn == 10) r = fast;                                      // the storages the call starts from, as they are in the state
      else if (fn == 11) r = slow;
      else {
        const double Q = lg_calc_CR_Q(cls.par, a[i], b[i], &fast, &slow);
        r = fn == 7 ? Q : fn == 8 ? fast : slow;
      }
    }
    out[i] = r;
  }
}
cudaError_t lgar_launch_debug_functions(const LgarConfig& cfg, const LgarDeviceState& d, int fn, int layer, int64_t n, const double* a, const double* b,
                                        double* out, cudaStream_t stream) {
  lgar_debug_functions_kernel<<<(unsigned)std::max<int64_t>(1, std::min<int64_t>((n + 127) / 128, 4096)), 128, 0, stream>>>(cfg, d, fn, layer, n, a, b, out);
  return cudaGetLastError();
}

// The 16 terms of lgar_global_mass_balance (reference src/lgar.cxx:1162-1216) of every column, [ncol][16] on the device
// (lgarb_get_global_mass_balance_device; same order and arithmetic as the host version in lgar_engine.cu).
__global__ void lgar_mass_balance_kernel(LgarDeviceState d, const double* __restrict__ volstart, const double* __restrict__ volchange, double* __restrict__ out) {
  const int64_t st = d.stride;
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < d.ncol; c += (int64_t)gridDim.x * blockDim.x) {
    double o[16];
    o[0] = volstart[c];
    o[1] = d.cum[LGAR_CUM_PRECIP * st + c];
    o[2] = d.cum[LGAR_CUM_IN * st + c];
    o[3] = d.scal[LGAR_S_STORAGE * st + c];
    o[4] = d.scal[LGAR_S_VOLON * st + c];
    o[5] = d.cum[LGAR_CUM_RUNOFF * st + c];
    o[6] = d.cum[LGAR_CUM_GIUH * st + c];
    o[7] = d.cum[LGAR_CUM_RECH * st + c];
    o[8] = d.cum[LGAR_CUM_AET * st + c];
    o[9] = d.cum[LGAR_CUM_PET * st + c];
    o[10] = d.cum[LGAR_CUM_RUNOFF_CR * st + c];
    o[11] = d.scal[LGAR_S_CR_FAST * st + c] + d.scal[LGAR_S_CR_SLOW * st + c];
    o[12] = d.cum[LGAR_CUM_Q * st + c];
    o[13] = volchange ? volchange[c] : 0.0;
    o[14] = d.cum[LGAR_CUM_Q_CR * st + c];
    o[15] = o[0] + o[1] - o[5] - o[8] - o[4] - o[7] - o[3] + o[13] - o[10] - o[11];  // lgar.cxx:1185
#pragma unroll
    for (int k = 0; k < 16; k++) out[c * 16 + k] = o[k];
  }
}
cudaError_t lgar_launch_mass_balance(const LgarDeviceState& d, const double* volstart, const double* volchange, double* out, cudaStream_t stream) {
  lgar_mass_balance_kernel<<<(unsigned)std::max<int64_t>(1, std::min<int64_t>((d.ncol + 255) / 256, 2048)), 256, 0, stream>>>(d, volstart, volchange, out);
  return cudaGetLastError();
}

// Gather per-front BMI arrays into a caller-facing [ncol][LGAR_W] layout (pad = NaN).
__global__ void lgar_gather_fronts_kernel(LgarDeviceState d, int which, double scale, double* dst, int64_t col0, int64_t ncol) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ncol * LGAR_W) return;
  const int64_t c = i / LGAR_W;
  const int slot = (int)(i % LGAR_W);
  const int64_t col = col0 + c;
  const double* src = which == 0 ? d.depth : which == 1 ? d.theta : which == 2 ? d.psi : which == 3 ? d.K : d.dzdt;
  dst[i] = (slot < d.nfront[col]) ? src[(int64_t)slot * d.stride + col] * scale : LG_NAN;
}

__global__ void lgar_gather_layers_kernel(LgarDeviceState d, double* dst, int64_t col0, int64_t ncol, int L) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ncol * L) return;
  const int64_t c = i / L;
  const int k = (int)(i % L);
  const LgarColumnClass* cls = d.classes + d.class_id[col0 + c];
  dst[i] = (k < cls->num_layers) ? cls->cum[k] : LG_NAN;  // quirk Q10: from index 0, in cm
}

// ---- launch wrappers ------------------------------------------------------------------------------
cudaError_t lgar_launch_step_stream(const LgarConfig& cfg, const LgarDeviceState& d, const LgarSinks& sinks, int force_all_active, cudaStream_t stream) {
  cudaError_t err = cudaMemsetAsync(d.work_count, 0, 2 * sizeof(int32_t), stream);
  if (err != cudaSuccess) return err;
  lgar_step_stream_kernel<<<(unsigned)((d.ncol + 255) / 256), 256, 0, stream>>>(cfg, d, sinks, force_all_active);
  return cudaGetLastError();
}

cudaError_t lgar_launch_step(const LgarConfig& cfg, const LgarDeviceState& d, const LgarSinks& sinks, int active_blocks,
                             int force_all_active, cudaStream_t stream, cudaEvent_t* ev) {
  cudaError_t err = cudaMemsetAsync(d.work_count, 0, 2 * sizeof(int32_t), stream);
  if (err != cudaSuccess) return err;
  const int threads = 256;
  const int64_t blocks = (d.ncol + threads - 1) / threads;
  if (ev) cudaEventRecord(ev[0], stream);
  lgar_step_stream_kernel<<<(unsigned)blocks, threads, 0, stream>>>(cfg, d, sinks, force_all_active);
  if (ev) cudaEventRecord(ev[1], stream);
  lgar_step_active_kernel<<<active_blocks, LGAR_ACTIVE_THREADS, 0, stream>>>(cfg, d, sinks);
  if (ev) cudaEventRecord(ev[2], stream);
  return cudaGetLastError();
}

__global__ void lgar_sum_fronts_kernel(LgarDeviceState d, unsigned long long* acc) {
  unsigned long long s = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < d.ncol; i += (int64_t)gridDim.x * blockDim.x) s += (unsigned)d.nfront[i];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xFFFFFFFFu, s, o);
  if ((threadIdx.x & 31) == 0 && s) atomicAdd(acc, s);
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(acc + 1, (unsigned long long)d.work_count[0]);
}

cudaError_t lgar_launch_sum_fronts(const LgarDeviceState& d, unsigned long long* acc, cudaStream_t stream) {
  lgar_sum_fronts_kernel<<<296, 256, 0, stream>>>(d, acc);
  return cudaGetLastError();
}

cudaError_t lgar_launch_gather_fronts(const LgarDeviceState& d, int which, double scale, double* dst, int64_t col0, int64_t ncol,
                                      cudaStream_t stream) {
  const int64_t n = ncol * LGAR_W;
  lgar_gather_fronts_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d, which, scale, dst, col0, ncol);
  return cudaGetLastError();
}

cudaError_t lgar_launch_gather_layers(const LgarDeviceState& d, double* dst, int64_t col0, int64_t ncol, int L, cudaStream_t stream) {
  const int64_t n = ncol * L;
  lgar_gather_layers_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d, dst, col0, ncol, L);
  return cudaGetLastError();
}

cudaError_t lgar_launch_window(const LgarConfig& cfg, const LgarDeviceState& d, const LgarWindow& w, int blocks, cudaStream_t stream, int mode) {
  cudaError_t err = cudaMemsetAsync(w.tile_counter, 0, sizeof(int32_t), stream);
  if (err != cudaSuccess) return err;
  (void)mode;  // mode 1 only: the per-lane staged kernel (mode 2, 6e7 column-timesteps/s) was removed in round 2
  lgar_window_kernel<<<blocks, LGAR_WINDOW_THREADS, 0, stream>>>(cfg, d, w);
  return cudaGetLastError();
}

int lgar_window_kernel_max_blocks_per_sm() {
  int nb = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, lgar_window_kernel, LGAR_WINDOW_THREADS, 0);
  return nb;
}

int lgar_active_kernel_max_blocks_per_sm() {
  int nb = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, lgar_step_active_kernel, LGAR_ACTIVE_THREADS, 0);
  return nb;
}
