// lgar_c_b200/csrc/lgar_finish_smem.cu -- finishing kernel of the wavefront scheduler with the lane records in SHARED memory
// (opt-in: LGARB_FINISH_SMEM=1; default off until it has been measured on a B200).
//
// lgar_wave_finish_kernel (lgar_kernels.cu) keeps the record of its slot in a 3.3 KB thread-private variable.  Local memory
// is interleaved by lane -- a 128-byte line holds one word of each of the 32 lanes of a warp -- so a warp's records occupy
// 32 x 3.3 KB = 104 KB of L1 however many of its lanes own a slot, 420 KB per block: the records live in L2 and every access
// of the serial chain that sets the length of that kernel pays an L2 round trip.  Here a block
// keeps its 128 / lane_stride records in dynamic shared memory (pitch 103 x 32 bytes: 32-byte aligned for the record's
// alignas(32) members, and an odd number of 32-byte units so that the owners of a warp's slots fall on different banks,
// two-way at worst).  The stage functions and their arithmetic are those of the kernel in lgar_kernels.cu; results are bit
// for bit the same (tests/test_gpu_parity.py::test_scheduling_invariance runs this variant beside the others).
//
// A translation unit of its own on purpose: the device functions of lgar_lanes.cuh are inlined per module by cost, and one
// more caller inside lgar_kernels.cu changed the SASS of every measured kernel there.
#define LGAR_TU_LOCAL 1
#include <cuda_runtime.h>

#include <algorithm>

#include "lgar_ctx.cuh"
#include "lgar_kernels.h"

constexpr int kFinishSmemPitch = (((int)sizeof(LgLaneSlot) / 32) | 1) * 32;
static_assert(kFinishSmemPitch >= (int)sizeof(LgLaneSlot) && kFinishSmemPitch % 64 == 32, "shared pitch: holds a record, odd multiple of 32 bytes");

__global__ void __launch_bounds__(128) lgar_wave_finish_smem_kernel(const __grid_constant__ LgarConfig cfg, const __grid_constant__ LgarDeviceState d, const __grid_constant__ LgarWindow w, const __grid_constant__ LgarWave wv, int lane_stride) {
  extern __shared__ __align__(128) unsigned char lg_finish_smem[];
  LgLaneSlot* slots = reinterpret_cast<LgLaneSlot*>(wv.ctx);
  int total = 0;
  for (int q = 0; q < 2 * LGAR_NSTAGE; q++) total += wv.qcount[q];
  unsigned long long front_sum = 0, active_sum = 0;
  unsigned long long* fsp = w.prof ? &front_sum : nullptr;
  if (threadIdx.x % lane_stride != 0) return;  // lane_stride divides the block size (launcher); no block-wide barrier below
  LgLaneSlot& S = *reinterpret_cast<LgLaneSlot*>(lg_finish_smem + (size_t)(threadIdx.x / lane_stride) * kFinishSmemPitch);
  // persistent blocks (the launcher sizes the grid to what is resident at once); a thread that has finished its column takes
  // the next item from the cursor, so a block is not held up by the order the items happen to sit in
  for (int i = atomicAdd(wv.cursor, 1); i < total; i = atomicAdd(wv.cursor, 1)) {
    int j = i, slot = -1, stage = LG_ST_DONE;
    for (int q = 0; q < 2 * LGAR_NSTAGE; q++) {  // item i of the concatenation of all queues
      const int n = wv.qcount[q];
      if (j < n) {
        slot = wv.queue[q >> 1][q & 1][j];
        stage = q >> 1;
        break;
      }
      j -= n;
    }
    if (slot < 0) continue;
    lane_load<LG_PART_ALL>(&S, &slots[slot]);
    LgLane& L = S.L;
    L.stage = stage;  // the queue a slot sits in is authoritative: the partial-context stages do not rewrite L.stage
    L.c.cfg = &cfg;
    L.c.cls = L.cls;
    L.c.f = &L.f;
    L.c.ff = lg_ff_ones;
    while (L.stage != LG_ST_DONE) {
      switch (L.stage) {
        case LG_ST_FETCH: lg_stage_fetch(cfg, d, w, L, fsp); break;
        case LG_ST_HEAD:
          lg_stage_head(cfg, d, w, L, fsp);
          active_sum += (L.stage == LG_ST_FRONT) ? 1 : 0;
          break;
        case LG_ST_FRONT: lg_stage_front(L); break;
        case LG_ST_SEARCH: {
          LgSearch q = L.q;  // the 176-byte search record in registers / thread-private memory for the trips
          while (!q.done) lg_search_iter(q, *L.cls);
          L.q = q;
          L.resume = true;
          L.stage = LG_ST_FRONT;
          break;
        }
        case LG_ST_TAIL: lg_stage_tail(cfg, d, w, L, fsp); break;
        case LG_ST_CORR:
          while (!L.k.done) lg_corr_iter(L.k, L.f, *L.cls);
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

// Converged variant with the records in shared memory (LGARB_FINISH_SMEM=2): one warp per block, every lane owns a record
// (32 x 3 296 B = 103 KB, two blocks per SM), the warp executes ONE stage at a time for the lanes that are in it and goes
// round the stages in program order -- lgar_wave_finish_vote_kernel of lgar_kernels.cu without the thread-private records --
// and a lane whose column has finished the window takes the next item from the cursor, so the warp stays full until the
// items run out.  With LGARB_WAVE_FINISH_BELOW above the column count this kernel runs the WHOLE window on chip: it is the
// measurement of what DESIGN.md section 10 item 1 proposes (step context on chip, ~2-4 warps per SM, no context traffic at
// stage boundaries), at the price of today's 32-front record size.
__global__ void __launch_bounds__(32) lgar_wave_finish_vote_smem_kernel(const __grid_constant__ LgarConfig cfg, const __grid_constant__ LgarDeviceState d, const __grid_constant__ LgarWindow w, const __grid_constant__ LgarWave wv, int quantum) {
  extern __shared__ __align__(128) unsigned char lg_finish_smem[];
  const unsigned FULL = 0xFFFFFFFFu;
  LgLaneSlot* slots = reinterpret_cast<LgLaneSlot*>(wv.ctx);
  int total = 0;
  for (int q = 0; q < 2 * LGAR_NSTAGE; q++) total += wv.qcount[q];
  unsigned long long front_sum = 0, active_sum = 0;
  unsigned long long* fsp = w.prof ? &front_sum : nullptr;
  LgLaneSlot& S = *reinterpret_cast<LgLaneSlot*>(lg_finish_smem + (size_t)threadIdx.x * kFinishSmemPitch);  // blockDim.x == 32
  LgLane& L = S.L;
  L.stage = LG_ST_DONE;
  bool exhausted = false;
  int cur = LGAR_NSTAGE - 1;
  for (;;) {
    if (L.stage == LG_ST_DONE && !exhausted) {  // refill: the next item of the concatenation of all queues
      int j = atomicAdd(wv.cursor, 1);
      if (j < total) {
        int slot = -1, stage = LG_ST_DONE;
        for (int q = 0; q < 2 * LGAR_NSTAGE; q++) {
          const int n = wv.qcount[q];
          if (j < n) {
            slot = wv.queue[q >> 1][q & 1][j];
            stage = q >> 1;
            break;
          }
          j -= n;
        }
        if (slot >= 0) {
          lane_load<LG_PART_ALL>(&S, &slots[slot]);
          L.c.cfg = &cfg;
          L.c.cls = L.cls;
          L.c.f = &L.f;
          L.c.ff = lg_ff_ones;
          L.stage = stage;  // the queue a slot sits in is authoritative
        }
      } else {
        exhausted = true;
      }
    }
    __syncwarp();
    unsigned m[LGAR_NSTAGE];
    unsigned any = 0;
#pragma unroll
    for (int s = 0; s < LGAR_NSTAGE; s++) {
      m[s] = __ballot_sync(FULL, L.stage == s);
      any |= m[s];
    }
    if (!any) {
      if (__all_sync(FULL, exhausted)) break;
      continue;  // some lane has just been told there is nothing left, or refills next turn
    }
#pragma unroll 1
    for (int k = 0; k < LGAR_NSTAGE; k++) {  // next non-empty stage after the one just run
      cur = (cur + 1 == LGAR_NSTAGE) ? 0 : cur + 1;
      if (m[cur]) break;
    }
    if (L.stage == cur) {
      switch (cur) {
        case LG_ST_FETCH: lg_stage_fetch(cfg, d, w, L, fsp); break;
        case LG_ST_HEAD:
          lg_stage_head(cfg, d, w, L, fsp);
          active_sum += (L.stage == LG_ST_FRONT) ? 1 : 0;
          break;
        case LG_ST_FRONT: lg_stage_front(L); break;
        case LG_ST_SEARCH: {
          LgSearch q = L.q;  // the search record in registers / thread-private memory for the trips
          for (int it = 0; it < quantum && !q.done; it++) lg_search_iter(q, *L.cls);
          L.q = q;
          if (q.done) {
            L.resume = true;
            L.stage = LG_ST_FRONT;
          }
          break;
        }
        case LG_ST_TAIL: lg_stage_tail(cfg, d, w, L, fsp); break;
        default:
          for (int it = 0; it < quantum && !L.k.done; it++) lg_corr_iter(L.k, L.f, *L.cls);
          if (L.k.done) L.stage = LG_ST_TAIL;
          break;
      }
    }
    __syncwarp();
  }
  if (w.prof) {
    if (front_sum) atomicAdd(w.prof, front_sum);
    if (active_sum) atomicAdd(w.prof + 1, active_sum);
  }
}

static __global__ void lgar_wave_clear_smem_kernel(LgarWave wv) {
  if (threadIdx.x < 2 * LGAR_NSTAGE) wv.qcount[threadIdx.x] = 0;
  if (threadIdx.x == 0) *wv.cursor = 0;  // zero between launches (the FETCH kernel relies on it)
}

cudaError_t lgar_launch_wave_finish_smem(const LgarConfig& cfg, const LgarDeviceState& d, const LgarWindow& w, const LgarWave& wv, int in_flight,
                                         int lane_stride, cudaStream_t stream) {
  if (lane_stride <= 0) {  // converged variant: one warp per block, -lane_stride = search trips per turn (0: 128)
    const int bytes = 32 * kFinishSmemPitch;
    // per launch (once per window): the attribute belongs to the current device, and engines may live on several
    cudaError_t rc0 = cudaFuncSetAttribute(lgar_wave_finish_vote_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (rc0 != cudaSuccess) return rc0;
    const int resident = std::max(1, (227 * 1024) / (bytes + 1024));
    const int blocks = std::max(1, std::min((in_flight + 31) / 32, 148 * resident));
    cudaError_t rc = cudaMemsetAsync(wv.cursor, 0, sizeof(int32_t), stream);
    if (rc != cudaSuccess) return rc;
    lgar_wave_finish_vote_smem_kernel<<<blocks, 32, bytes, stream>>>(cfg, d, w, wv, lane_stride < 0 ? -lane_stride : 128);
    lgar_wave_clear_smem_kernel<<<1, 32, 0, stream>>>(wv);
    return cudaGetLastError();
  }
  int ls = 2;  // a power of two in [2, 32]: 64 records (206 KB) per block at most
  while (ls < lane_stride && ls < 32) ls *= 2;
  const int per_block = 128 / ls;
  const int bytes = per_block * kFinishSmemPitch;
  cudaError_t rc0 = cudaFuncSetAttribute(lgar_wave_finish_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (rc0 != cudaSuccess) return rc0;
  const int resident = std::max(1, (227 * 1024) / (bytes + 1024));  // blocks per SM that fit the shared memory
  const int blocks = std::max(1, std::min((in_flight + per_block - 1) / per_block, 148 * resident));
  cudaError_t rc = cudaMemsetAsync(wv.cursor, 0, sizeof(int32_t), stream);
  if (rc != cudaSuccess) return rc;
  lgar_wave_finish_smem_kernel<<<blocks, 128, bytes, stream>>>(cfg, d, w, wv, ls);
  lgar_wave_clear_smem_kernel<<<1, 32, 0, stream>>>(wv);
  return cudaGetLastError();
}
