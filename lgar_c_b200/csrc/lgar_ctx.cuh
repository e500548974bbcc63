// lgar_c_b200/csrc/lgar_ctx.cuh -- the lane record's way between HBM and a stage's own copy: queue append, 256-bit
// context copies, the parts of the record each stage moves.  Shared by lgar_kernels.cu (stage kernels, finishing kernels)
// and lgar_finish_smem.cu (finishing kernel with the records in shared memory).
#pragma once
#include <cstddef>

#include "lgar_lanes.cuh"

namespace {

// bins of the stage queues: by front count (and a flag), by the layer of the searched front
__device__ __forceinline__ int wave_bin_fronts(int n, int flags = 0) {
#if LGAR_NBIN_N >= 6
  int b = n <= 2 ? 0 : n == 3 ? 1 : n == 4 ? 2 : n == 5 ? 3 : n <= 7 ? 4 : 5;
#else
  int b = n <= 3 ? 0 : n == 4 ? 1 : n <= 6 ? 2 : 3;
#endif
  if (b >= LGAR_NBIN_N) b = LGAR_NBIN_N - 1;
  return b + LGAR_NBIN_N * (flags & (LGAR_NBIN_FLAG - 1));   // LGAR_NBIN_FLAG is 1, 2 or 4
}
// flags of a slot on its way to TAIL: bit 0 the step creates a surficial front, bit 1 the front list needs a correction
__device__ __forceinline__ int wave_tail_flags(const LgLane& L) { return (L.create ? 1 : 0) | ((L.me_phase == 1 || L.me_phase == 2) && L.me_correction != 0 ? 2 : 0); }
__device__ __forceinline__ int wave_bin_layers(int l) { return l <= 1 ? 0 : l >= LGAR_NBIN ? LGAR_NBIN - 1 : l - 1; }

// Append `slot` to bin `bin` of the queue of stage `next` (next outside [0, LGAR_NSTAGE): nothing to append).  One atomicAdd per
// distinct (stage, bin) among the lanes of the warp that call it together.
__device__ __forceinline__ void wave_push(const LgarWave& wv, unsigned append_mask, int next, int bin, int slot) {
  const bool pred = next >= 0 && next < LGAR_NSTAGE;
  const unsigned act = __ballot_sync(__activemask(), pred);
  if (!pred) return;
  const int buf = (append_mask >> next) & 1;
  const int key = (next * 2 + buf) * LGAR_NBIN + bin;
  const unsigned peers = __match_any_sync(act, key);
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(peers) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(wv.qcount + key, __popc(peers));
  base = __shfl_sync(peers, base, leader);
  wv.queue[next][buf][(size_t)bin * wv.nslots + base + __popc(peers & ((1u << lane) - 1))] = slot;
}

// The index space a consumer of queue (stage, buf) walks: the bins one after the other, each starting at a multiple of 32.  The
// table lives in shared memory (one per block, filled by its first warp): with 24 bins it would cost every thread 48 registers.
struct WaveItems {
  int cnt[LGAR_NBIN], voff[LGAR_NBIN + 1];
};
__device__ __forceinline__ void wave_items_init(WaveItems& it, const LgarWave& wv, int stage, int buf) {
  if (threadIdx.x == 0) {
    it.voff[0] = 0;
    for (int b = 0; b < LGAR_NBIN; b++) {
      it.cnt[b] = wv.qcount[(stage * 2 + buf) * LGAR_NBIN + b];
      it.voff[b + 1] = it.voff[b] + ((it.cnt[b] + 31) & ~31);
    }
  }
  __syncthreads();
}
// the slot at index i of that space, or -1 (padding at the end of a bin)
__device__ __forceinline__ int wave_item(const LgarWave& wv, const WaveItems& it, int stage, int buf, int i) {
  int b = 0;
#pragma unroll 1
  while (b + 1 < LGAR_NBIN && i >= it.voff[b + 1]) b++;
  const int j = i - it.voff[b];
  return j < it.cnt[b] ? wv.queue[stage][buf][(size_t)b * wv.nslots + j] : -1;
}
// End of a consumer of queue (stage, consume): the last of its blocks to get here clears the counts of that queue (and the FETCH
// cursor), so that the queue is empty when its turn to be appended to comes -- round 1 launched a one-thread kernel per stage launch
// for this (half of all launches).  Every thread of the block must call it.  Nobody reads those counts any more: every block read
// them when it started, and it is the last block that clears them.
__device__ __forceinline__ void wave_consumed(const LgarWave& wv, int stage, int consume, bool reset_cursor) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(wv.done_blocks + stage, 1) + 1 == (int)gridDim.x) {
      for (int b = 0; b < LGAR_NBIN; b++) wv.qcount[(stage * 2 + consume) * LGAR_NBIN + b] = 0;
      if (reset_cursor) *wv.cursor = 0;
      wv.done_blocks[stage] = 0;
      __threadfence();
    }
  }
}

// item j of the concatenation of every bin of every queue (the finishing kernels): slot and the stage it is parked in
__device__ __forceinline__ int wave_item_any(const LgarWave& wv, int j, int* stage) {
  for (int q = 0; q < 2 * LGAR_NSTAGE * LGAR_NBIN; q++) {
    const int n = wv.qcount[q];
    if (j < n) {
      *stage = q / (2 * LGAR_NBIN);
      return wv.queue[q / (2 * LGAR_NBIN)][(q / LGAR_NBIN) & 1][(size_t)(q % LGAR_NBIN) * wv.nslots + j];
    }
    j -= n;
  }
  return -1;
}

// Context copy between HBM and thread-private memory.  The HBM side moves 32 bytes per lane per instruction
// (LDG/STG.E.256, new on sm_100): a lane's record is contiguous, so every access is one full 32-byte sector --
// with 16-byte accesses each sector is requested twice (ncu: lts write sectors = 2x the bytes stored).
// NBYTES = leading part of the record to move.
template <int NBYTES, typename T>
__device__ __forceinline__ void ctx_load(T* local_dst, const T* global_src) {
  static_assert(NBYTES % 32 == 0 && NBYTES <= (int)sizeof(T), "context part must be a multiple of 32 bytes");
  const char* g = reinterpret_cast<const char*>(global_src);
  ulonglong2* l = reinterpret_cast<ulonglong2*>(local_dst);
#pragma unroll 8
  for (int i = 0; i < NBYTES / 32; i++) {
    unsigned long long a, b, c, e;
    asm volatile("ld.global.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(e) : "l"(g + 32 * i));
    l[2 * i] = make_ulonglong2(a, b);
    l[2 * i + 1] = make_ulonglong2(c, e);
  }
}
template <int NBYTES, typename T>
__device__ __forceinline__ void ctx_store(T* global_dst, const T* local_src) {
  static_assert(NBYTES % 32 == 0 && NBYTES <= (int)sizeof(T), "context part must be a multiple of 32 bytes");
  char* g = reinterpret_cast<char*>(global_dst);
  const ulonglong2* l = reinterpret_cast<const ulonglong2*>(local_src);
#pragma unroll 8
  for (int i = 0; i < NBYTES / 32; i++) {
    const ulonglong2 x = l[2 * i], y = l[2 * i + 1];
    asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(g + 32 * i), "l"(x.x), "l"(x.y), "l"(y.x), "l"(y.y) : "memory");
  }
}

// dynamic-length variants (nchunks 32-byte chunks)
__device__ __forceinline__ void ctx_load_n(char* l, const char* g, int nchunks) {
#pragma unroll 1
  for (int i = 0; i < nchunks; i++) {
    unsigned long long a, b, c, e;
    asm volatile("ld.global.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(e) : "l"(g + 32 * i));
    reinterpret_cast<ulonglong2*>(l)[2 * i] = make_ulonglong2(a, b);
    reinterpret_cast<ulonglong2*>(l)[2 * i + 1] = make_ulonglong2(c, e);
  }
}
__device__ __forceinline__ void ctx_store_n(char* g, const char* l, int nchunks) {
#pragma unroll 1
  for (int i = 0; i < nchunks; i++) {
    const ulonglong2 x = reinterpret_cast<const ulonglong2*>(l)[2 * i], y = reinterpret_cast<const ulonglong2*>(l)[2 * i + 1];
    asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(g + 32 * i), "l"(x.x), "l"(x.y), "l"(y.x), "l"(y.y) : "memory");
  }
}

}  // namespace

struct alignas(32) LgLaneSlot {
  LgLane L;
};

// Parts of the lane record (lgar_lanes.cuh) a stage moves between HBM and its thread-private copy.  The front
// arrays have capacity LGAR_W = 32 but a column carries 4-5 fronts on average: only the live prefix of each
// array travels (counts kept in the record's first sector), which removes about a third of the context traffic.
// Every stage moves only what it reads or writes: HEAD leaves the epilogue state (GIUH queue, cumulative volumes) where it is,
// FRONT moves the fronts, 128 bytes of step context, state_previous (in only) and the search record -- not the scalars, not the
// rest of the step context.
enum {
  LG_PART_CTX = 1,       // LgCtx (status word; the pointers are set by every stage)
  LG_PART_S = 2,         // scalars, frozen factors
  LG_PART_EP = 4,        // epilogue state
  LG_PART_FRONTS = 8,
  LG_PART_STEPF = 16,    // the front loop's part of the step context
  LG_PART_STEPR = 32,    // the rest of it, up to and including the correction record
  LG_PART_OLD = 64,
  LG_PART_SEARCH = 128,
  LG_PART_STATE = LG_PART_CTX | LG_PART_S | LG_PART_EP | LG_PART_FRONTS,
  LG_PART_STEP = LG_PART_STEPF | LG_PART_STEPR,
  LG_PART_ALL = 255
};
constexpr int kOffC = (int)offsetof(LgLane, c), kOffS = (int)offsetof(LgLane, s), kOffEp = (int)offsetof(LgLane, ep), kOffF = (int)offsetof(LgLane, f),
              kOffPro = (int)offsetof(LgLane, pro), kOffStepR = (int)offsetof(LgLane, P), kOffOld = (int)offsetof(LgLane, old), kOffQ = (int)offsetof(LgLane, q);
constexpr int kFrontTail = (int)offsetof(LgFronts, layer);  // layer[], tb[], n behind the five double arrays
static_assert(kOffC == 32 && kOffS % 32 == 0 && kOffEp % 32 == 0 && kOffF % 32 == 0 && kOffPro % 32 == 0 && kOffStepR == kOffPro + 128 && kOffOld % 32 == 0 &&
                  kOffQ % 32 == 0 && sizeof(LgLaneSlot) % 32 == 0,
              "LgLane layout");
constexpr int kPitch = 8 * LGAR_W;                           // bytes between the arrays of LgFronts / LgOld
constexpr int kFrontTailBytes = kOffPro - kOffF - kFrontTail;  // layer[], tb[], n (+ padding up to `pro`)
static_assert(offsetof(LgFronts, depth) == 0 && offsetof(LgFronts, theta) == kPitch && offsetof(LgFronts, psi) == 2 * kPitch &&
                  offsetof(LgFronts, K) == 3 * kPitch && offsetof(LgFronts, dzdt) == 4 * kPitch && kFrontTail == 5 * kPitch &&
                  kFrontTailBytes % 32 == 0 && kFrontTailBytes >= 2 * LGAR_W + 4 && kPitch % 32 == 0,
              "LgFronts layout");
static_assert(offsetof(LgOld, theta) == kPitch && offsetof(LgOld, psi) == 2 * kPitch && offsetof(LgOld, n) == 3 * kPitch && kOffQ - kOffOld == 3 * kPitch + 32,
              "LgOld layout");
static_assert(offsetof(LgLane, nold_stored) < 32, "front counts must sit in the first sector of the record");

#define LG_CTX_RANGE(fn, from, to) fn<(to) - (from)>(reinterpret_cast<LgLaneSlot*>(l + (from)), reinterpret_cast<const LgLaneSlot*>(g + (from)))

template <int PARTS>
__device__ __forceinline__ void lane_load(LgLaneSlot* S, const LgLaneSlot* G) {
  char* l = reinterpret_cast<char*>(S);
  const char* g = reinterpret_cast<const char*>(G);
  ctx_load<32>(S, G);  // first sector: stage, hour, column, class, front counts
  const int cf = (S->L.nf_stored + 3) >> 2, co = (S->L.nold_stored + 3) >> 2;
  if (PARTS & LG_PART_CTX) LG_CTX_RANGE(ctx_load, kOffC, kOffS);
  if (PARTS & LG_PART_S) LG_CTX_RANGE(ctx_load, kOffS, kOffEp);
  if (PARTS & LG_PART_EP) LG_CTX_RANGE(ctx_load, kOffEp, kOffF);
  if (PARTS & LG_PART_FRONTS) {
#pragma unroll
    for (int a = 0; a < 5; a++) ctx_load_n(l + kOffF + kPitch * a, g + kOffF + kPitch * a, cf);
    LG_CTX_RANGE(ctx_load, kOffF + kFrontTail, kOffPro);
  }
  if (PARTS & LG_PART_STEPF) LG_CTX_RANGE(ctx_load, kOffPro, kOffStepR);
  if (PARTS & LG_PART_STEPR) LG_CTX_RANGE(ctx_load, kOffStepR, kOffOld);
  if (PARTS & LG_PART_OLD) {
#pragma unroll
    for (int a = 0; a < 3; a++) ctx_load_n(l + kOffOld + kPitch * a, g + kOffOld + kPitch * a, co);
    LG_CTX_RANGE(ctx_load, kOffOld + 3 * kPitch, kOffQ);
  }
  if (PARTS & LG_PART_SEARCH) LG_CTX_RANGE(ctx_load, kOffQ, (int)sizeof(LgLaneSlot));
#ifdef LGAR_POISON  // validation build: whatever was not moved must never be read
  if (PARTS & LG_PART_FRONTS)
    for (int i = 4 * cf; i < LGAR_W; i++) S->L.f.depth[i] = S->L.f.theta[i] = S->L.f.psi[i] = S->L.f.K[i] = S->L.f.dzdt[i] = LG_NAN;
  if (PARTS & LG_PART_OLD)
    for (int i = 4 * co; i < LGAR_W; i++) S->L.old.depth[i] = S->L.old.theta[i] = S->L.old.psi[i] = LG_NAN;
#endif
}
#undef LG_CTX_RANGE
#define LG_CTX_RANGE(fn, from, to) fn<(to) - (from)>(reinterpret_cast<LgLaneSlot*>(g + (from)), reinterpret_cast<const LgLaneSlot*>(l + (from)))
template <int PARTS>
__device__ __forceinline__ void lane_store(LgLaneSlot* G, LgLaneSlot* S) {
  const char* l = reinterpret_cast<const char*>(S);
  char* g = reinterpret_cast<char*>(G);
  if (PARTS & LG_PART_FRONTS) S->L.nf_stored = S->L.f.n;
  if (PARTS & LG_PART_OLD) S->L.nold_stored = S->L.old.n;
  const int cf = (S->L.nf_stored + 3) >> 2, co = (S->L.nold_stored + 3) >> 2;
  ctx_store<32>(G, S);
  if (PARTS & LG_PART_CTX) LG_CTX_RANGE(ctx_store, kOffC, kOffS);
  if (PARTS & LG_PART_S) LG_CTX_RANGE(ctx_store, kOffS, kOffEp);
  if (PARTS & LG_PART_EP) LG_CTX_RANGE(ctx_store, kOffEp, kOffF);
  if (PARTS & LG_PART_FRONTS) {
#pragma unroll
    for (int a = 0; a < 5; a++) ctx_store_n(g + kOffF + kPitch * a, l + kOffF + kPitch * a, cf);
    LG_CTX_RANGE(ctx_store, kOffF + kFrontTail, kOffPro);
  }
  if (PARTS & LG_PART_STEPF) LG_CTX_RANGE(ctx_store, kOffPro, kOffStepR);
  if (PARTS & LG_PART_STEPR) LG_CTX_RANGE(ctx_store, kOffStepR, kOffOld);
  if (PARTS & LG_PART_OLD) {
#pragma unroll
    for (int a = 0; a < 3; a++) ctx_store_n(g + kOffOld + kPitch * a, l + kOffOld + kPitch * a, co);
    LG_CTX_RANGE(ctx_store, kOffOld + 3 * kPitch, kOffQ);
  }
  if (PARTS & LG_PART_SEARCH) LG_CTX_RANGE(ctx_store, kOffQ, (int)sizeof(LgLaneSlot));
}
#undef LG_CTX_RANGE
