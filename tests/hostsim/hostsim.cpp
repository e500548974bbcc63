// tests/hostsim/hostsim.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Compiles the kernel text (lgar_c_b200/csrc/lgar_physics.cuh + lgar_step.cuh) with a host compiler
// and drives it for ONE column, so the kernel's logic can be compared with the oracle on a machine
// without a GPU.  It is never linked into, nor loaded by, liblgarb.so or the lgar_c_b200 package:
// the product has no CPU path.  Arithmetic differs from the GPU only in libm's pow vs CUDA's pow.
#define LGAR_HOSTSIM 1
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../lgar_c_b200/csrc/lgar_config.hpp"
#include "../../lgar_c_b200/csrc/lgar_lanes.cuh"

using namespace lgarhost;

// psi searches through the fast path of lgar_search.cuh (1, default) or the plain trip-by-trip walk (0)
static int hostsim_use_ff = 1;
extern "C" void hostsim_set_use_ff(int on) { hostsim_use_ff = on; }
// path counters (LGAR_PATH_* of lgar_types.h), accumulated over every run since the last reset
static unsigned long long hostsim_paths[LGAR_NPATH];
extern "C" void hostsim_get_paths(unsigned long long* out, int reset) {
  for (int i = 0; i < LGAR_NPATH; i++) out[i] = hostsim_paths[i];
  if (reset) memset(hostsim_paths, 0, sizeof(hostsim_paths));
}

#ifdef LGAR_FF_STATS
// how the psi searches went (searches, finished by the fast path, declined, fallbacks, evaluations fast/plain, trips, verify mismatches)
extern "C" void hostsim_ff_stats(long long* out8, int reset) {
  const long long v[8] = {lg_ff_stats.searches, lg_ff_stats.ff_done, lg_ff_stats.declined, lg_ff_stats.fallback, lg_ff_stats.evals_ff,
                          lg_ff_stats.evals_plain, lg_ff_stats.trips, lg_ff_stats.verify_bad};
  for (int i = 0; i < 8; i++) out8[i] = v[i];
  if (reset) lg_ff_stats = LgFfStats{};
}
#endif

extern "C" int hostsim_run_series(const char* cfg_path, const int* types_in, const double* thick_in, int nsteps, const double* P,
                                  const double* E, double* scalars, int* nfronts, int cap, double* fronts, int* flags,
                                  int force_all_active, int* status_out, char* err, int errcap) {
  try {
    ParsedConfig pc = parse_config(cfg_path);
    const int L = (int)pc.layer_thickness.size();
    if (L < 1 || L > LGAR_MAXL) throw std::runtime_error("unsupported number of layers");                                 // do_initialize, lgar_engine.cu
    if ((int)pc.layer_soil_type.size() != L) throw std::runtime_error("layer_soil_type and layer_thickness differ in length");
    std::vector<SoilRow> soils;
    const int nread = read_soil_file(pc.soil_params_file, pc.num_soil_types, pc.wilting_point_psi, pc.log_mode, soils, cfg_path);
    LgarConfig cfg = make_engine_config(pc);
    std::vector<int> types(L);
    std::vector<double> thick(L);
    for (int k = 0; k < L; k++) {
      types[k] = types_in ? types_in[k] : pc.layer_soil_type[k];
      thick[k] = thick_in ? thick_in[k] : pc.layer_thickness[k];
    }
    check_column_class(pc, nread, L, types.data(), thick.data());
    LgarColumnClass cls = make_class(pc, soils, L, types.data(), thick.data());

    const int64_t st = 1;
    std::vector<double> depth(LGAR_W), theta(LGAR_W), psi(LGAR_W), K(LGAR_W), dzdt(LGAR_W), scal(LGAR_NSCAL), giuhq(LGAR_MAXG), cum(LGAR_NCUM),
        out(LGAR_NOUT);
    unsigned long long fflags_w[LGAR_FW] = {0};  // stride 1: word w of the column at fflags_w[w]
    unsigned long long& fflags = fflags_w[0];
    int32_t nf = 0, bits = 1 << 8, status = 0, class_id = 0, work[1], work_count[2];
    double p1, e1;
    LgarDeviceState d;
    memset(&d, 0, sizeof(d));
    d.ncol = 1; d.stride = st;
    d.depth = depth.data(); d.theta = theta.data(); d.psi = psi.data(); d.K = K.data(); d.dzdt = dzdt.data();
    d.fflags = fflags_w; d.nfront = &nf; d.scal = scal.data(); d.bits = &bits; d.status = &status; d.class_id = &class_id;
    d.giuhq = giuhq.data(); d.cum = cum.data(); d.out = out.data(); d.classes = &cls; d.precip = &p1; d.pet = &e1;
    d.work = work; d.work_count = work_count;
    d.paths = hostsim_paths; d.use_ff = hostsim_use_ff;
    if (!cls.invalid_soil_type) {
      for (int k = 0; k < L; k++) {
        depth[k] = cls.cum[k + 1]; theta[k] = cls.initial_theta[k + 1]; psi[k] = pc.initial_psi; K[k] = cls.initial_K[k + 1]; dzdt[k] = 0.0;
        fflags |= (unsigned long long)(((k + 1) & 7) | 8) << (4 * k);
      }
      nf = L;
      scal[LGAR_S_STORAGE] = class_initial_storage(cls);
    }
    scal[LGAR_S_TIMESTEP_H] = pc.timestep_h;
    LgarSinks sinks;
    memset(&sinks, 0, sizeof(sinks));
    int t;
    if (force_all_active >= 2) {
      // window-kernel emulation (one-column tile): advance through cached hours, then one active step
      std::vector<double> sk((size_t)LGAR_NOUT * nsteps, 0.0);
      std::vector<int32_t> nfs((size_t)nsteps, 0);
      LgarWindow w;
      memset(&w, 0, sizeof(w));
      w.precip = P; w.pet = E; w.series_stride = 1; w.sink_stride = 1; w.K = nsteps; w.force_all_active = (force_all_active == 3);
      for (int k = 0; k < LGAR_NOUT; k++) w.sinks.p[k] = sk.data() + (size_t)k * nsteps;
      w.has_sinks = 1;
      w.use_ff = hostsim_use_ff;
      w.paths = hostsim_paths;
      w.nf_sink = nfs.data();
      if (force_all_active >= 4) {
        // lane-program emulation: one lane walks the stages of lgar_lanes.cuh
        w.force_all_active = (force_all_active == 5);
        int32_t counter = 0;
        w.tile_counter = &counter;
        static LgLane L;  // large: keep it off the stack
        L.stage = LG_ST_FETCH; L.col = -1; L.t = nsteps;
        while (L.stage != LG_ST_DONE) {
          switch (L.stage) {
            case LG_ST_FETCH: lg_stage_fetch(cfg, d, w, L, nullptr); break;
            case LG_ST_HEAD: lg_stage_head(cfg, d, w, L, nullptr); break;
            case LG_ST_FRONT: lg_stage_front(L); break;
            case LG_ST_SEARCH:
              lg_search_run(L.q, *L.cls, hostsim_use_ff != 0);
              L.resume = true; L.stage = LG_ST_FRONT; break;
            case LG_ST_TAIL: lg_stage_tail(cfg, d, w, L, nullptr); break;
            case LG_ST_CORR:
              lg_corr_run(L.k, L.f, *L.cls, hostsim_use_ff != 0);
              L.stage = LG_ST_TAIL; break;
          }
        }
      } else {
      int tt = 0;
      while (tt < nsteps) {
        tt = lg_advance_column(cfg, d, w, 0, tt, nullptr);
        if (tt < nsteps) { lg_active_column_at(cfg, d, w, 0, tt, nullptr); tt++; }
      }
      }
      for (t = 0; t < nsteps; t++) {
        if (scalars) for (int k = 0; k < LGAR_NOUT; k++) scalars[(size_t)t * LGAR_NOUT + k] = sk[(size_t)k * nsteps + t];
        if (nfronts) nfronts[t] = nfs[t];
      }
      if (fronts) {  // only the final front list is available in window mode: store it at the last step
        t = nsteps - 1;
        for (int i = 0; i < nf && i < cap; i++) {
          double* q = fronts + ((size_t)t * cap + i) * 5;
          q[0] = depth[i]; q[1] = theta[i]; q[2] = psi[i]; q[3] = K[i]; q[4] = dzdt[i];
          int nib = (int)((fflags_w[i >> 4] >> (4 * (i & 15))) & 0xF);
          flags[(size_t)t * cap + i] = (nib & 7) | ((nib >> 3) << 8);
        }
      }
      if (status_out) *status_out = status;
      return nsteps;
    }
    for (t = 0; t < nsteps; t++) {
      p1 = P[t]; e1 = E[t];
      if (lg_stream_column(cfg, d, sinks, 0, force_all_active)) lg_active_column(cfg, d, sinks, 0);
      if (status & LGAR_FAIL_MASK) break;
      if (scalars) for (int k = 0; k < LGAR_NOUT; k++) scalars[(size_t)t * LGAR_NOUT + k] = out[k];
      if (nfronts) nfronts[t] = nf;
      if (fronts) {
        for (int i = 0; i < nf && i < cap; i++) {
          double* q = fronts + ((size_t)t * cap + i) * 5;
          q[0] = depth[i]; q[1] = theta[i]; q[2] = psi[i]; q[3] = K[i]; q[4] = dzdt[i];
          int nib = (int)((fflags_w[i >> 4] >> (4 * (i & 15))) & 0xF);
          flags[(size_t)t * cap + i] = (nib & 7) | ((nib >> 3) << 8);
        }
      }
    }
    if (status_out) *status_out = status;
    return t;
  } catch (const std::exception& ex) {
    if (err && errcap > 0) snprintf(err, (size_t)errcap, "%s", ex.what());
    return -1;
  }
}
