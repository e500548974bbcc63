// lgar_c_b200/csrc/lgar_finish_smem.cu -- finishing kernel of the wavefront scheduler with the lane records in SHARED memory.
//
// When few slots are left in flight (the columns that compute fluxes in nearly every hour of the window) a round of stage
// launches costs ~2 ms whatever it holds, so the engine hands every remaining slot to ONE kernel that runs it to the end of the
// window.  Round 1's version (lgar_wave_finish_kernel, lgar_kernels.cu) gives each slot a thread that works out of a 3.3 KB
// thread-private record -- which lives in L2: 0.5 ms per active step of the busiest column.  Here a warp keeps the records of
// its slots in shared memory (a record per owner lane; slots_per_warp = 32 ... 2), executes ONE stage at a time for the lanes
// that are in it and goes round the stages in program order, so its lanes meet again in the psi search, which is where the time
// goes; a lane whose column has finished the window takes the next slot from the cursor.  No context traffic at stage
// boundaries, every access of the serial chain at shared-memory latency.
//
// Measured (profiles/r2_call5_call6_summary.md, 1.25 M columns): a call of 480 h 4.79e8 -> 4.90e8 column-timesteps/s with 32 records
// per warp, 5.00e8 with 16; a call of 1 460 h 5.29e8 -> 5.48e8.  The same kernel fed EARLY with the busiest columns, beside the
// rounds (an "express lane": slots that advance fewer than two hours per round leave the rounds from round 24 on), was measured
// too and removed: 3.3e8 -- its shared memory takes the L1 away from the stage kernels, whose bulk rounds went from 13.4 to 22 ms
// while it ran, and 296 warps with 32 diverged lanes each sustain only 1.7e7 steps/s against 9.3e7 for the rounds.
//
// A translation unit of its own: the device functions of lgar_lanes.cuh are inlined per module by cost, and one more caller
// inside lgar_kernels.cu would change the code of every stage kernel there.  The arithmetic is the stage functions' own:
// tests/test_gpu_scale_parity.py::test_scheduling_does_not_change_results runs it beside round 1's kernel, bit for bit.
#define LGAR_TU_LOCAL 1
#include <cuda_runtime.h>

#include <algorithm>

#include "lgar_ctx.cuh"
#include "lgar_kernels.h"

namespace {

constexpr int kXPitch = (((int)sizeof(LgLaneSlot) / 32) | 1) * 32;
static_assert(kXPitch >= (int)sizeof(LgLaneSlot) && kXPitch % 64 == 32, "shared pitch: holds a record, odd multiple of 32 bytes");

template <bool FF>
__global__ void __launch_bounds__(32) lgar_wave_finish_smem_kernel(const __grid_constant__ LgarConfig cfg, const __grid_constant__ LgarDeviceState d,
                                                                   const __grid_constant__ LgarWindow w, const __grid_constant__ LgarWave wv,
                                                                   int lane_stride, int quantum) {
  extern __shared__ __align__(128) unsigned char lg_xsmem[];
  const unsigned FULL = 0xFFFFFFFFu;
  const int lane = threadIdx.x;
  LgLaneSlot* slots = reinterpret_cast<LgLaneSlot*>(wv.ctx);
  int total = 0;
  for (int q = 0; q < 2 * LGAR_NSTAGE * LGAR_NBIN; q++) total += wv.qcount[q];
  unsigned long long front_sum = 0, active_sum = 0;
  unsigned long long* fsp = w.prof ? &front_sum : nullptr;
  const bool owner = (lane % lane_stride) == 0;  // lane_stride 1 ... 16: 32 ... 2 records per warp
  LgLaneSlot& S = *reinterpret_cast<LgLaneSlot*>(lg_xsmem + (size_t)(lane / lane_stride) * kXPitch);
  LgLane& L = S.L;
  int stage = LG_ST_DONE;     // the lane's stage lives in a register: L.stage is written by the stage functions
  bool exhausted = !owner;
  int cur = LGAR_NSTAGE - 1;
  for (;;) {
    if (stage == LG_ST_DONE && !exhausted) {  // refill
      int j = atomicAdd(wv.cursor, 1);
      if (j < total) {
        int st = LG_ST_DONE;
        const int slot = wave_item_any(wv, j, &st);  // item j of the concatenation of all stage queues; the queue a slot sits in is its stage
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
  for (int i = threadIdx.x; i < 2 * LGAR_NSTAGE * LGAR_NBIN; i += blockDim.x) wv.qcount[i] = 0;
  if (threadIdx.x == 0) *wv.cursor = 0;  // zero between launches (the FETCH kernel relies on it)
}

}  // namespace

cudaError_t lgar_launch_wave_finish_smem(const LgarConfig& cfg, const LgarDeviceState& d, const LgarWindow& w, const LgarWave& wv, int in_flight,
                                         int slots_per_warp, int quantum, int use_ff, int sm_count, cudaStream_t stream) {
  int spw = 32;
  while (spw > 2 && spw > slots_per_warp) spw >>= 1;  // 32, 16, 8, 4 or 2 records per warp
  const int bytes = spw * kXPitch;
  // largest dynamic size opted in to so far, per instantiation, on the device it was set on
  static int configured[2] = {-1, -1};
  static int configured_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev != configured_dev) {
    configured[0] = configured[1] = -1;
    configured_dev = dev;
  }
  const int which = use_ff ? 1 : 0;
  if (bytes > configured[which]) {
    cudaError_t rc = use_ff ? cudaFuncSetAttribute(lgar_wave_finish_smem_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)
                            : cudaFuncSetAttribute(lgar_wave_finish_smem_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (rc != cudaSuccess) return rc;
    configured[which] = bytes;
  }
  const int resident = std::max(1, std::min(32, (227 * 1024) / (bytes + 1024)));  // blocks per SM that fit the shared memory (32: the block limit)
  const int blocks = std::max(1, std::min((in_flight + spw - 1) / spw, sm_count * resident));
  if (use_ff)
    lgar_wave_finish_smem_kernel<true><<<blocks, 32, bytes, stream>>>(cfg, d, w, wv, 32 / spw, quantum);
  else
    lgar_wave_finish_smem_kernel<false><<<blocks, 32, bytes, stream>>>(cfg, d, w, wv, 32 / spw, quantum);
  lgar_wave_xclear_kernel<<<1, 256, 0, stream>>>(wv);
  return cudaGetLastError();
}
