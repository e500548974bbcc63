// lgar_c_b200/csrc/lgar_express.cu -- the express lane of the wavefront scheduler: whole columns on chip.
//
// What it is for.  The stage kernels of lgar_kernels.cu advance every slot in flight by (about) one active step per round.
// A column that computes fluxes in nearly every hour (a wet profile under a wet climate: 1-2 % of the bench workload) therefore
// needs as many ROUNDS as the window has hours, while the average column needs a seventh of that: once the bulk has finished,
// the rounds run nearly empty (latency bound, 1.5-2 ms whatever they hold) and the window ends with a kernel that carries the
// stragglers one thread each -- out of a 3.3 KB thread-private record that lives in L2 (0.5 ms per step, measured).  The longer
// the call, the larger that share: one year in calls of 1 460 h ran at 2.9e8 column-timesteps/s against 4.7e8 for a call of 480 h
// (profiles/r2_call5_summary.md).
//
// Here a warp keeps the lane records of up to 32 such columns in SHARED memory (a record per lane, pitch = an odd number of
// 32-byte units so that neighbouring owners fall on different banks) and runs them to the end of the window: no context traffic
// at stage boundaries, every access of the serial chain at shared-memory latency.  The warp executes ONE stage at a time for the
// lanes that are in it and goes round the stages in program order, so its lanes meet again in the psi search, which is where
// the time goes; a lane whose column has finished the window takes the next entry from the cursor.
//
// Two uses (lgar_engine.cu, run_wave_window):
//   * express lane: from round `express_r0` on, FETCH hands slots that have advanced fewer than `factor` hours per round to the
//     express queue instead of the HEAD queue; the engine launches this kernel for them on streams of its own, beside the rounds.
//   * finishing kernel: when few slots are left in flight, every remaining slot (whatever stage it is parked in).
//
// A translation unit of its own: the device functions of lgar_lanes.cuh are inlined per module by cost, and one more caller
// inside lgar_kernels.cu would change the code of every stage kernel there.  The arithmetic is the stage functions' own:
// tests/test_gpu_scale_parity.py::test_scheduling_does_not_change_results runs both uses beside the plain rounds, bit for bit.
#define LGAR_TU_LOCAL 1
#include <cuda_runtime.h>

#include <algorithm>

#include "lgar_ctx.cuh"
#include "lgar_kernels.h"

namespace {

constexpr int kXPitch = (((int)sizeof(LgLaneSlot) / 32) | 1) * 32;
static_assert(kXPitch >= (int)sizeof(LgLaneSlot) && kXPitch % 64 == 32, "shared pitch: holds a record, odd multiple of 32 bytes");

template <bool FF>
__global__ void __launch_bounds__(32) lgar_wave_express_kernel(const __grid_constant__ LgarConfig cfg, const __grid_constant__ LgarDeviceState d,
                                                               const __grid_constant__ LgarWindow w, const __grid_constant__ LgarWave wv, int xbegin,
                                                               int xend, int32_t* cursor, int lane_stride, int quantum) {
  extern __shared__ __align__(128) unsigned char lg_xsmem[];
  const unsigned FULL = 0xFFFFFFFFu;
  const int lane = threadIdx.x;
  const bool finishing = xend < 0;
  LgLaneSlot* slots = reinterpret_cast<LgLaneSlot*>(wv.ctx);
  int total = 0;
  if (finishing)
    for (int q = 0; q < 2 * LGAR_NSTAGE; q++) total += wv.qcount[q];
  else
    total = xend - xbegin;
  unsigned long long front_sum = 0, active_sum = 0;
  unsigned long long* fsp = w.prof ? &front_sum : nullptr;
  const bool owner = (lane % lane_stride) == 0;  // lane_stride 1, 2 or 4: 32, 16 or 8 records per warp
  LgLaneSlot& S = *reinterpret_cast<LgLaneSlot*>(lg_xsmem + (size_t)(lane / lane_stride) * kXPitch);
  LgLane& L = S.L;
  int stage = LG_ST_DONE;     // the lane's stage lives in a register: L.stage is written by the stage functions
  bool exhausted = !owner;
  int cur = LGAR_NSTAGE - 1;
  for (;;) {
    if (stage == LG_ST_DONE && !exhausted) {  // refill
      int j = atomicAdd(cursor, 1);
      if (j < total) {
        int slot = -1, st = LG_ST_DONE;
        if (finishing) {  // item j of the concatenation of all stage queues; the queue a slot sits in is its stage
          for (int q = 0; q < 2 * LGAR_NSTAGE; q++) {
            const int n = wv.qcount[q];
            if (j < n) {
              slot = wv.queue[q >> 1][q & 1][j];
              st = q >> 1;
              break;
            }
            j -= n;
          }
        } else {
          slot = wv.xqueue[xbegin + j];
          st = LG_ST_HEAD;
        }
        if (slot >= 0) {
          lane_load<LG_PART_ALL>(&S, &slots[slot]);
          L.c.cfg = &cfg;
          L.c.cls = L.cls;
          L.c.f = &L.f;
          L.c.ff = L.ffv;
          L.c.paths = w.paths;
          L.c.use_ff = false;
          L.stage = st;
          stage = st;
        }
      } else {
        exhausted = true;
      }
    }
    unsigned m[LGAR_NSTAGE];
    unsigned any = 0;
#pragma unroll
    for (int s = 0; s < LGAR_NSTAGE; s++) {
      m[s] = __ballot_sync(FULL, stage == s);
      any |= m[s];
    }
    if (!any) {
      if (__all_sync(FULL, exhausted)) break;
      continue;  // a lane has just learnt that nothing is left, the others refill next turn
    }
#pragma unroll 1
    for (int k = 0; k < LGAR_NSTAGE; k++) {  // the next non-empty stage after the one that has just run
      cur = (cur + 1 == LGAR_NSTAGE) ? 0 : cur + 1;
      if (m[cur]) break;
    }
    if (stage == cur) {
      switch (cur) {
        case LG_ST_FETCH: lg_stage_fetch(cfg, d, w, L, fsp); break;
        case LG_ST_HEAD:
          lg_stage_head(cfg, d, w, L, fsp);
          active_sum += (L.stage == LG_ST_FRONT) ? 1 : 0;
          break;
        case LG_ST_FRONT: lg_stage_front(L); break;
        case LG_ST_SEARCH: {
          LgSearch q = L.q;  // the search record in registers for the trips
          lg_search_some(q, *L.cls, quantum, FF);
          L.q = q;
          if (q.done) {
            L.resume = true;
            L.stage = LG_ST_FRONT;
          }
          break;
        }
        case LG_ST_TAIL: lg_stage_tail(cfg, d, w, L, fsp); break;
        default:
          lg_corr_some(L.k, L.f, *L.cls, quantum, FF);
          if (L.k.done) L.stage = LG_ST_TAIL;
          break;
      }
      stage = L.stage;
    }
    __syncwarp();
  }
  if (w.prof) {
    if (front_sum) atomicAdd(w.prof, front_sum);
    if (active_sum) atomicAdd(w.prof + 1, active_sum);
  }
}

__global__ void lgar_wave_xclear_kernel(LgarWave wv) {
  if (threadIdx.x < 2 * LGAR_NSTAGE) wv.qcount[threadIdx.x] = 0;
  if (threadIdx.x == 0) *wv.cursor = 0;  // zero between launches (the FETCH kernel relies on it)
}

}  // namespace

cudaError_t lgar_launch_wave_express(const LgarConfig& cfg, const LgarDeviceState& d, const LgarWindow& w, const LgarWave& wv, int xbegin, int xend,
                                     int32_t* cursor, int items, int slots_per_warp, int quantum, int use_ff, int sm_count, cudaStream_t stream) {
  const int spw = slots_per_warp >= 32 ? 32 : slots_per_warp >= 16 ? 16 : 8;
  const int bytes = spw * kXPitch;
  static int configured[2] = {-1, -1};  // largest dynamic size opted in to so far, per instantiation (set on the current device)
  int dev = 0;
  cudaGetDevice(&dev);
  static int configured_dev = -1;
  if (dev != configured_dev) {
    configured[0] = configured[1] = -1;
    configured_dev = dev;
  }
  const int which = use_ff ? 1 : 0;
  if (bytes > configured[which]) {
    cudaError_t rc = use_ff ? cudaFuncSetAttribute(lgar_wave_express_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)
                            : cudaFuncSetAttribute(lgar_wave_express_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (rc != cudaSuccess) return rc;
    configured[which] = bytes;
  }
  const int resident = std::max(1, std::min(32, (227 * 1024) / (bytes + 1024)));  // blocks per SM that fit the shared memory
  const int blocks = std::max(1, std::min((items + spw - 1) / spw, sm_count * resident));
  cudaError_t rc = cudaMemsetAsync(cursor, 0, sizeof(int32_t), stream);
  if (rc != cudaSuccess) return rc;
  if (use_ff)
    lgar_wave_express_kernel<true><<<blocks, 32, bytes, stream>>>(cfg, d, w, wv, xbegin, xend, cursor, 32 / spw, quantum);
  else
    lgar_wave_express_kernel<false><<<blocks, 32, bytes, stream>>>(cfg, d, w, wv, xbegin, xend, cursor, 32 / spw, quantum);
  if (xend < 0) lgar_wave_xclear_kernel<<<1, 32, 0, stream>>>(wv);
  return cudaGetLastError();
}
