// lgar_c_b200/csrc/lgar_lanes.cuh -- the Update() of a column as a resumable per-lane program.
//
// Why: the psi search of lgar_theta_mass_balance (reference src/lgar.cxx:2952-3112) is ~75 % of the
// FP64 work and its trip count is heavy tailed (median 7 theta evaluations per active step, mean
// 115, max > 20 000).  Run as ordinary nested calls, 32 columns in lock step keep 2.5 of 32 lanes
// busy (ncu: smsp__thread_inst_executed_per_inst_executed = 2.47, profiles/).  Here every lane runs
// its own stream of (column, hour) steps and is always in exactly one STAGE; the warp executes one
// stage at a time for all lanes that are in it, so lanes meet again in the search loop no matter
// where in their column's step they are:
//
//   FETCH   take a column (atomic counter), walk it through the hours that need no flux computation
//   HEAD    start of an active step: load state, sub-step / AET / free-drainage set-up, start the
//           front loop of lgar_move_wetting_fronts (src/lgar.cxx:1313)
//   FRONT   iterations of that loop (deepest front first) until one needs a psi search
//   SEARCH  iterations of the psi search, all searching lanes converged
//   TAIL    merge/cross corrections, dzdt, reservoir, mass balance; next sub-cycle or end of step
//   CORR    iterations of the correction search of lgar_fix_dry_over_wet_wetting_fronts, converged
//
// The arithmetic and its order are those of lgar_physics.cuh (the two are checked against each
// other bit for bit, tests/test_hostsim.py and tests/test_gpu_parity.py); only the control
// structure differs.
#ifndef LGAR_LANES_CUH
#define LGAR_LANES_CUH

#include "lgar_step.cuh"


// The lane record.  Its head is the column's STATE (scalars, GIUH queue, cumulative volumes, front list): it is
// imported from the SoA arrays when the lane takes the column and exported when the column has finished its
// window (lg_lane_import/export), so between those two points no stage touches the SoA state -- the wavefront
// kernels move contiguous records, not 45 scattered 8-byte values per step.  Behind it the context of the step in
// flight.  Field order matters: the wavefront kernels copy only the leading part a stage needs (state: up to
// `pro`; TAIL: up to `old`), see lgar_wave_kernel.
struct LgLane {
  int stage;
  int t;
  int64_t col;
  const LgarColumnClass* cls;
  int nf_stored, nold_stored;  // front counts the record was last stored with: only that prefix of the arrays is moved
  LgCtx c;
  alignas(32) LgScalars s;
  double ffv[LGAR_MAXL + 2];  // frozen_factor per layer (1-based) of this column for the window: all 1.0 unless the run is SFT-coupled
  alignas(32) LgEpi ep;
  alignas(32) LgFronts f;
  // ---- step context (locals of lg_step_active) ----
  // first the 128 bytes the front loop works on (FRONT moves these and nothing else of the step context, lgar_ctx.cuh)
  alignas(32) LgPrologue pro;
  int wf, nwf, wf_fd_demand, search_case;   // front-loop context (locals of lg_move_wetting_fronts)
  int me_phase, me_correction;              // resumable lgar_move_wetting_fronts epilogue (the correction type is found at the end of the front loop)
  bool create, resume;
  double precip_mass_to_add, bottom_flux, actual_ET_demand, AET, free_drainage, mass_corr_fd, volend_sub;
  // then the rest
  alignas(32) double P;
  double E;
  LgStepVols v;
  LgCycle cy;
  int cycle, wf_fd;
  double PET_rate, ponded, volon_ts, temp_rch;
  double volin, runoff, volon_sub, last_AET, last_PET_sub, last_volrech;
  // lgar_fix_dry_over_wet_wetting_fronts, resumable around its correction search
  int fx_cur, fx_nx, fx_l;
  double fx_prior_mass, fx_mass_change;
  bool fx_resume;
  unsigned int tr_iters, tr_kcyc;  // work trace of the current column (debug)
  alignas(16) LgCorr k;   // TAIL/CORR stop copying after this
  alignas(32) LgOld old;  // FRONT only
  alignas(32) LgSearch q; // FRONT/SEARCH only
};

// column state: SoA arrays <-> lane record
LG_FN_NOINLINE void lg_lane_import(const LgarConfig& cfg, const LgarDeviceState& d, int64_t col, LgLane& L) {
  L.cls = d.classes + d.class_id[col];
  load_scalars(d, col, L.s);
  load_fronts(d, col, L.f);
  lg_epi_load(cfg, d, col, L.ep);
  L.nf_stored = L.f.n;
  L.nold_stored = 0;
  // SFT coupling: the column's frozen factors stay what they are for the whole window (lgarb_set_value("soil_temperature_profile")
  // between two calls changes them for the next one)
  for (int k = 0; k < LGAR_MAXL + 2; k++) L.ffv[k] = (d.ff && k >= 1 && k <= LGAR_MAXL) ? d.ff[(int64_t)k * d.stride + col] : 1.0;
}
LG_FN_NOINLINE void lg_lane_export(const LgarConfig& cfg, const LgarDeviceState& d, int64_t col, const LgLane& L) {
  store_scalars(d, col, L.s, true);
  store_fronts(d, col, L.f);
  lg_epi_store(cfg, d, col, L.ep);
}

// ---- psi search split into init / iterate / finish (lgar.cxx:2952-3112) ---------------------------
// returns true when the search has to iterate (otherwise q.theta is final)
LG_FN bool lg_search_init(LgLane& L, int layer_num, double psi_cm, double new_mass, double prior_mass) {
  LgSearch& q = L.q;
  lg_search_setup(q, layer_num, psi_cm, new_mass, prior_mass, L.precip_mass_to_add);
  if (q.delta_mass <= LG_MBAL_TOL) {
    q.theta = lg_theta_from_h(psi_cm, L.cls->lay[layer_num]);
    q.done = true;
    return false;
  }
  return true;
}

// what follows the loop in lgar_theta_mass_balance (lgar.cxx:3096-3110); only after an iterated search
LG_FN void lg_search_finish(LgLane& L) {
  const LgSearch& q = L.q;
  if (q.iter == 0) return;
  lg_path_search_exit(L.c, q);
  if ((q.delta_mass > LG_MBAL_TOL) && (!q.wanted_to_saturate)) L.AET = L.AET - fabs(q.delta_mass - LG_MBAL_TOL);
  if ((q.theta >= L.cls->lay[q.layer_num].theta_e) && (q.psi_loc != 0.0)) L.AET = 0.0;
}

// ---- lgar_move_wetting_fronts as begin / front iterations / end (lgar.cxx:1269-1866) --------------
LG_FN void lg_move_begin(LgLane& L, double volin_arg) {
  L.nwf = L.f.n;
  L.wf = L.nwf;
  L.precip_mass_to_add = volin_arg;
  L.bottom_flux = 0.0;
  L.wf_fd_demand = L.wf_fd;
  L.resume = false;
  L.me_phase = 0;
}

// the saturated-front depth adjustment executed when wf == 1 (lgar.cxx:1644-1754)
LG_FN_NOINLINE void lg_move_top_block(LgLane& L) {
  LgCtx& c = L.c;
  LgFronts& f = L.f;
  const LgarColumnClass& cls = *L.cls;
  const double* cum = cls.cum;
  const int num_layers = cls.num_layers;
  const double old_mass = L.volend_sub, free_drainage_cm = L.free_drainage, mass_corr_fd = L.mass_corr_fd;
  const double precip_mass_to_add = L.precip_mass_to_add;
  double actual_ET_demand = L.actual_ET_demand;
  L.wf_fd_demand = lg_wf_free_drainage(f);
  const int fdi = L.wf_fd_demand - 1;
  const double theta_e_k1 = cls.lay[f.layer[fdi]].theta_e;
  double mass_timestep = (old_mass + precip_mass_to_add) - (actual_ET_demand + free_drainage_cm + mass_corr_fd);
  if (!(old_mass > 0.0)) c.status |= LGAR_FAIL_OLD_MASS;
  if (fabs(f.theta[fdi] - theta_e_k1) < 1E-15) {
    double current_mass = lg_mass_bal(c);
    double mass_balance_error = fabs(current_mass - mass_timestep);
    if (fabs(mass_balance_error) > LG_MBAL_TOL) LG_PATH(c, LGAR_PATH_SATURATED_DEPTH);
    double factor = 1.0;
    if (fdi + 1 < f.n) {
      if (fabs(f.theta[fdi] - f.theta[fdi + 1]) < 0.01) factor = factor / fabs(f.theta[fdi] - f.theta[fdi + 1]);
    }
    bool switched = false, aug = false, break_flag = false;
    double depth_new = f.depth[fdi];
    int iter = 0;
    const int speedup_thresh = LG_MAX_ITER_SAT / 10;
    while (fabs(mass_balance_error) > LG_MBAL_TOL) {
      iter++;
      if (iter > LG_MAX_ITER_SAT) {
        break_flag = true;
        L.AET = L.AET + mass_balance_error;
        actual_ET_demand = L.AET;
        break;
      }
      if ((iter > speedup_thresh) && (!aug)) {
        factor = factor * 100;
        aug = true;
      }
      if (current_mass < mass_timestep) {
        depth_new += 0.01 * factor;
        switched = false;
      } else {
        if (!switched) {
          switched = true;
          factor = factor * 0.1;
        }
        depth_new -= 0.01 * factor;
      }
      if (f.tb[fdi] && (f.layer[fdi] == num_layers)) depth_new = cum[num_layers];
      f.depth[fdi] = depth_new;
      current_mass = lg_mass_bal(c);
      mass_balance_error = fabs(current_mass - mass_timestep);
    }
    if (depth_new < LG_TRUNCATION_DEPTH) {
      depth_new = LG_TRUNCATION_DEPTH;
      f.depth[fdi] = depth_new;
    }
    if (isinf(f.depth[fdi])) {
      f.depth[fdi] = cum[num_layers] + 1.E-6;
      break_flag = true;
    }
    if (break_flag) {
      current_mass = lg_mass_bal(c);
      mass_timestep = (old_mass + precip_mass_to_add) - (actual_ET_demand + free_drainage_cm + mass_corr_fd);
      L.bottom_flux += mass_timestep - current_mass;
    }
  }
  L.actual_ET_demand = actual_ET_demand;
}

// Iterations of the front loop until a search must iterate (returns true, L.q is armed) or the
// loop is over (returns false).  On entry with L.resume the search of front L.wf has just finished.
LG_FN_NOINLINE bool lg_move_fronts(LgLane& L) {
  LgCtx& c = L.c;
  LgFronts& f = L.f;
  const LgOld& old = L.old;
  const LgarColumnClass& cls = *L.cls;
  const double* cum = cls.cum;
  const int num_layers = cls.num_layers;
  const double column_depth = cum[num_layers];
  const int nwf = L.nwf;
  const double free_drainage_cm = L.free_drainage, mass_corr_fd = L.mass_corr_fd, timestep_h = L.pro.subtimestep_h;

  while (L.wf != 0) {
    const int wf = L.wf;
    const int ci = wf - 1, ni = wf;
    if (ci >= f.n || ci >= old.n) {
      c.status |= LGAR_FAIL_LIST;
      L.wf = 0;
      return false;
    }
    int layer_num = f.layer[ci];
    const LgarLayer* spp = &cls.lay[layer_num];
    if (!L.resume) {
      int layer_below;
      if (wf == nwf)
        layer_below = layer_num + 1;
      else {
        if (ni >= f.n) {
          c.status |= LGAR_FAIL_LIST;
          L.wf = 0;
          return false;
        }
        layer_below = f.layer[ni];
      }
      L.actual_ET_demand = L.AET;
      L.search_case = 0;

      if ((wf < nwf) && (layer_below != layer_num)) {  // :1379-1387
        f.theta[ci] = lg_theta_from_h(f.psi[ni], *spp);
        f.psi[ci] = f.psi[ni];
      }

      if (wf == nwf && layer_below != layer_num && nwf == num_layers) {  // :1403-1480
        f.depth[ci] += f.dzdt[ci] * timestep_h;
        double psi_old = old.psi[ci], psi_cm = f.psi[ci];
        double prior_mass = (old.depth[ci] - cum[layer_num - 1]) * (old.theta[ci] - 0.0);
        double new_mass = (f.depth[ci] - cum[layer_num - 1]) * (f.theta[ci] - 0.0);
        for (int k = 1; k < layer_num; k++) {
          const LgarLayer& sk = cls.lay[k];
          double theta_old = lg_theta_from_h(psi_old, sk);
          prior_mass += (sk.thick * theta_old);
          double theta = lg_theta_from_h(psi_cm, sk);
          new_mass += sk.thick * theta;
          L.q.dth[k] = 0.0;
          L.q.dz[k] = sk.thick;
        }
        L.q.dth[layer_num] = 0.0;
        L.q.dz[layer_num] = f.depth[ci] - cum[layer_num - 1];
        if (L.wf_fd_demand == wf) prior_mass += L.precip_mass_to_add - (free_drainage_cm + mass_corr_fd + L.actual_ET_demand);
        L.search_case = 2;
        if (lg_search_init(L, layer_num, psi_cm, new_mass, prior_mass)) return true;
      } else if ((wf < nwf) && (layer_num == layer_below)) {  // :1489-1637
        if (ni >= old.n) {
          c.status |= LGAR_FAIL_LIST;
          L.wf = 0;
          return false;
        }
        if (layer_num == 1) {
          const LgarLayer& sp = *spp;
          double prior_mass = old.depth[ci] * (old.theta[ci] - old.theta[ni]);
          if (L.wf_fd_demand == wf) prior_mass += L.precip_mass_to_add - (free_drainage_cm + mass_corr_fd + L.actual_ET_demand);
          f.depth[ci] += f.dzdt[ci] * timestep_h;
          if (f.depth[ci] > column_depth) f.depth[ci] = column_depth + LG_TRUNCATION_DEPTH;
          if (f.dzdt[ci] == 0.0 && !f.tb[ci]) {
            // a front that was just created keeps its theta
          } else {
            double th = prior_mass / f.depth[ci] + f.theta[ni];
            if (th < sp.theta_r) {  // :1519-1537
              LG_PATH(c, LGAR_PATH_THETA_R_DELETE);
              double mass_before = lg_mass_bal(c) - f.depth[ci] * (f.theta[ci] - th);
              lg_front_delete(c, ci);
              double mass_after = lg_mass_bal(c);
              L.AET = L.AET - fabs(mass_before - mass_after);
              L.actual_ET_demand = L.AET;
            } else {
              f.theta[ci] = fmax(sp.theta_r, fmin(sp.theta_e, th));
            }
          }
          f.psi[ci] = lg_h_from_Se(lg_Se_from_theta(f.theta[ci], sp), sp);  // :1634-1635
        } else {
          f.depth[ci] += f.dzdt[ci] * timestep_h;
          if (f.depth[ci] > column_depth) f.depth[ci] = column_depth + LG_TRUNCATION_DEPTH;
          double psi_old = old.psi[ci], psi_below_old = old.psi[ni];
          double psi_cm = f.psi[ci], psi_below = f.psi[ni];
          double prior_mass = (old.depth[ci] - cum[layer_num - 1]) * (old.theta[ci] - old.theta[ni]);
          double new_mass = (f.depth[ci] - cum[layer_num - 1]) * (f.theta[ci] - f.theta[ni]);
          for (int k = 1; k < layer_num; k++) {
            const LgarLayer& sk = cls.lay[k];
            double theta_old = lg_theta_from_h(psi_old, sk);
            double theta_below_old = lg_theta_from_h(psi_below_old, sk);
            prior_mass += (sk.thick * (theta_old - theta_below_old));
            double theta = lg_theta_from_h(psi_cm, sk);
            double theta_below = lg_theta_from_h(psi_below, sk);
            new_mass += sk.thick * (theta - theta_below);
            L.q.dth[k] = theta_below;
            L.q.dz[k] = sk.thick;
          }
          L.q.dth[layer_num] = f.theta[ni];
          L.q.dz[layer_num] = f.depth[ci] - cum[layer_num - 1];
          if (L.wf_fd_demand == wf) prior_mass += L.precip_mass_to_add - (free_drainage_cm + mass_corr_fd + L.actual_ET_demand);
          L.search_case = 3;
          if (lg_search_init(L, layer_num, psi_cm, new_mass, prior_mass)) return true;
        }
      }
    }
    // the search of this front (if any) is complete
    L.resume = false;
    if (L.search_case == 2 || L.search_case == 3) {
      const LgarLayer& sp = *spp;
      lg_search_finish(L);
      L.actual_ET_demand = L.AET;
      f.theta[ci] = fmax(sp.theta_r, fmin(L.q.theta, sp.theta_e));
      f.psi[ci] = lg_h_from_Se(lg_Se_from_theta(f.theta[ci], sp), sp);
      L.search_case = 0;
    }
    if (wf == 1) lg_move_top_block(L);
    L.wf = wf - 1;
  }
  return false;
}

// ---- correction search split into init / iterate (lgar.cxx:3262-3435) ---------------------------------
// returns true when the loop has to iterate
LG_FN_NOINLINE bool lg_corr_init(LgLane& L, int front_num, double prior_mass) {
  LgCtx& c = L.c;
  LgFronts& f = L.f;
  LgCorr& k = L.k;
  int cur = front_num - 1;
  k.done = true;
  k.iter = 0;
  if (cur < 0 || cur >= f.n) {
    c.status |= LGAR_FAIL_LIST;
    return false;
  }
  if (cur + 1 > 1) {
    if (f.tb[cur - 1]) cur = cur - 1;
  }
  while (f.tb[cur]) {
    if (cur + 1 == 1) break;
    cur = cur - 1;
    if (!f.tb[cur]) {
      cur = cur + 1;
      break;
    }
  }
  k.cur = cur;
  k.new_mass = lg_mass_bal(c);
  const double psi_cm = f.psi[cur];
  k.psi_loc = psi_cm;
  k.prior_mass = prior_mass;
  k.delta_mass = fabs(k.new_mass - prior_mass);
  k.factor = fmax(1.0, psi_cm / 10.0);
  if (psi_cm > 1.E4) k.factor = psi_cm;
  k.switched = false;
  k.psi_prev = psi_cm;
  k.delta_mass_prev = k.delta_mass;
  k.count_no_mass_change = 0;
  k.aug = false;
  k.aug_extreme = false;
  k.no_ff = false;
  k.done = !(k.delta_mass > LG_MBAL_TOL);
  if (!k.done) LG_PATH(c, LGAR_PATH_CORR_SEARCH);
  return !k.done;
}

// lgar_fix_dry_over_wet_wetting_fronts (lgar.cxx:2260-2307), resumable around its correction search.
// Returns true when a correction search has to iterate.
LG_FN_NOINLINE bool lg_fix_run(LgLane& L) {
  LgCtx& c = L.c;
  LgFronts& f = L.f;
  while (L.fx_l <= f.n) {
    if (L.fx_nx >= 0) {
      bool fixed = L.fx_resume;
      if (!L.fx_resume) {
        if ((f.theta[L.fx_cur] <= f.theta[L.fx_nx]) && (f.layer[L.fx_cur] == f.layer[L.fx_nx])) {
          L.fx_prior_mass = lg_mass_bal(c);
          L.fx_cur = lg_front_delete(c, L.fx_cur);
          fixed = true;
          if (lg_corr_init(L, L.fx_cur + 1, L.fx_prior_mass)) {
            L.fx_resume = true;
            return true;
          }
        }
      }
      L.fx_resume = false;
      if (fixed) {
        double mass_after = lg_mass_bal(c);
        L.fx_mass_change += (mass_after - L.fx_prior_mass);
      }
      L.fx_cur = L.fx_cur + 1;
      if (L.fx_cur >= f.n)
        L.fx_nx = -1;
      else
        L.fx_nx = (L.fx_cur + 1 < f.n) ? L.fx_cur + 1 : -1;
    }
    L.fx_l++;
  }
  return false;
}

// corrections + refresh after the front loop (lgar.cxx:1777-1863), resumable.  Returns true when a
// correction search has to iterate; otherwise the bottom boundary flux is in L.bottom_flux.
LG_FN_NOINLINE bool lg_move_end(LgLane& L) {
  LgCtx& c = L.c;
  LgFronts& f = L.f;
  const LgarColumnClass& cls = *L.cls;
  const unsigned lanes_here = LG_LANES_HERE();
  bool need_search = false;
  if (L.me_phase == 0) {
    L.me_correction = lg_correction_type(c);
    L.me_phase = 1;
  }
  while (L.me_correction != 0) {
    if (c.status & (LGAR_FAIL_LIST | LGAR_FAIL_OVERFLOW)) break;
    if (L.me_phase == 1) {
      const int correction = L.me_correction;
      LG_PATH(c, correction == 1 ? LGAR_PATH_MERGE : correction == 2 ? LGAR_PATH_CROSS_LAYER : correction == 3 ? LGAR_PATH_CROSS_BOTTOM : LGAR_PATH_DRY_OVER_WET);
      if (correction == 1) lg_merge(c);
      if (correction == 2) {
        double mass_before = lg_mass_bal(c);
        lg_cross_layer_boundary(c);
        double mass_after = lg_mass_bal(c);
        if (fabs(mass_before - mass_after) > 100. * LG_MBAL_TOL) L.AET = L.AET + (mass_before - mass_after);
      }
      if (correction == 3) {
        L.bottom_flux += lg_cross_domain_boundary(c);
        if (isnan(L.bottom_flux)) L.bottom_flux = 0.0;
      }
      if (correction == 4) {
        L.fx_mass_change = 0.0;
        L.fx_cur = 0;
        L.fx_nx = (f.n > 1) ? 1 : -1;
        L.fx_l = 1;
        L.fx_resume = false;
        L.me_phase = 2;
      }
    }
    if (L.me_phase == 2) {
      if (lg_fix_run(L)) {
        need_search = true;  // leave through the common exit: see LG_RECONVERGE
        break;
      }
      L.AET = L.AET - L.fx_mass_change;
      L.me_phase = 1;
    }
    L.me_correction = lg_correction_type(c);
  }
  LG_RECONVERGE(lanes_here);  // the few lanes that had corrections to make rejoin the others for the refresh
  if (!need_search) {
    for (int i = 0; i < f.n; i++) {
      const LgarLayer& sk = cls.lay[f.layer[i]];
      double Se = lg_Se_from_theta(f.theta[i], sk);
      if (f.psi[i] > 1.0) f.psi[i] = lg_h_from_Se(Se, sk);
      f.K[i] = lg_K_from_Se(Se, sk, c.ff[f.layer[i]] * sk.Ksat);  // lgar.cxx:1836
    }
    if (f.n == 1) {
      if (f.depth[0] != cls.cum[1]) f.depth[0] = cls.cum[1];
    }
    L.me_phase = 3;  // done
  }
  return need_search;
}

// ---- lg_step_active as cycle_begin / cycle_end ------------------------------------------------------
// Start of a sub-cycle up to (and including the set-up of) the front loop (bmi_lgar.cxx:356-652).
LG_FN_NOINLINE void lg_cycle_begin(const LgarConfig& cfg, LgLane& L) {
  LgCtx& c = L.c;
  LgFronts& f = L.f;
  LgScalars& s = L.s;
  const LgarColumnClass& cls = *L.cls;
  const double dt = L.pro.subtimestep_h;
  LgCycle& cy = L.cy;
  lg_cycle_head(cfg, s, cy, L.P, L.E, dt, L.volon_ts, L.PET_rate, L.ponded);
  LG_PATH(c, LGAR_PATH_ACTIVE_STEP);
  L.old.n = f.n;
  for (int i = 0; i < f.n; i++) {
    L.old.depth[i] = f.depth[i];
    L.old.theta[i] = f.theta[i];
    L.old.psi[i] = f.psi[i];
  }
  cy.volstart = lg_mass_bal(c);
  const double mass_check = cy.volstart;
  const int wf_fd = lg_wf_free_drainage(f);
  L.wf_fd = wf_fd;
  const double min_water_fd = lg_min_water_fd(c, wf_fd);
  const double storage_fd = lg_storage_fd(c, wf_fd);
  double mass_corr_fd = 0.0, free_drainage = 0.0;
  L.temp_rch = 0.0;
  if (L.pro.switch_caching) {
    L.PET_rate += s.acc_pet;
    s.acc_pet = 0.0;
    mass_corr_fd += s.acc_fd;
    s.acc_fd = 0.0;
  }
  if (cfg.free_drainage_enabled) {
    const int last = f.n - 1;
    free_drainage += dt * f.K[last];
    int it = 0;
    while ((mass_check - free_drainage < cls.min_storage) || (storage_fd - free_drainage < min_water_fd)) {
      free_drainage *= 0.5;
      if (it > 5) {
        free_drainage = 0.0;
        break;
      }
      it++;
    }
    if (free_drainage < 1.E-7) free_drainage = 0.0;
    if (f.psi[last] > 1.E6) free_drainage = 0.0;
  }
  const double precip_prev_sub = s.precip_prev;
  double AET = 0.0;
  if (L.PET_rate > 0.0) AET = lg_calc_aet(c, L.PET_rate, dt);
  {
    int it = 0;
    while ((mass_check - AET < cls.min_storage) || (storage_fd - AET < min_water_fd)) {
      AET *= 0.5;
      if (it > 5) {
        AET = 0.0;
        break;
      }
      it++;
    }
    it = 0;
    while (((mass_check - AET - free_drainage - mass_corr_fd) < cls.min_storage) ||
           ((storage_fd - AET - free_drainage - mass_corr_fd) < min_water_fd)) {
      AET *= 0.5;
      free_drainage *= 0.5;
      mass_corr_fd *= 0.5;
      if (it > 5) {
        AET = 0.0;
        free_drainage = 0.0;
        mass_corr_fd = 0.0;
        break;
      }
      it++;
    }
  }
  L.v.precip += cy.precip_sub + cy.precip_for_CR * dt;
  L.v.PET += fmax(cy.PET_sub, 0.0);
  L.volon_ts = fmax(L.volon_ts, 0.0);
  L.volon_ts = L.volon_ts > LG_SMALL_EPS ? L.volon_ts : 0.0;
  const double theta_e_top = cls.lay[f.layer[0]].theta_e;
  const bool top_saturated = (f.theta[0] + LG_SMALL_EPS) >= theta_e_top;
  cy.top_near_sat = f.theta[0] > theta_e_top * cls.par.spf_factor;
  bool create = (precip_prev_sub == 0.0 && cy.precip_sub > 0.0 && L.volon_ts == 0) ||
                ((cy.precip_sub > 0.0 || L.volon_ts > 0) && (f.n == cls.num_layers));
  if (top_saturated || L.volon_ts > 0.0) create = false;
  L.create = create;
  L.AET = AET;
  L.free_drainage = free_drainage;
  L.mass_corr_fd = mass_corr_fd;
  L.volin = 0.0;
  L.runoff = 0.0;
  L.volon_sub = 0.0;
  if (create) {
    lg_move_begin(L, 0.0);  // temp_pd = 0 (:549-556)
  } else {
    if (L.ponded > 0) {  // :601-618
      L.runoff = lg_insert_water(c, dt, AET, free_drainage, &L.ponded, &L.volin, cy.precip_rate, wf_fd);
      L.volon_sub = L.ponded;
      if (L.runoff < 0) c.status |= LGAR_FAIL_NEG_RUNOFF;
    } else {  // :619-632
      if (L.ponded < cls.par.ponded_depth_max_cm) {
        L.volon_sub = L.ponded;
        L.ponded = 0.0;
        L.runoff = 0.0;
      } else {
        L.runoff = (L.ponded - cls.par.ponded_depth_max_cm);
        L.volon_sub = cls.par.ponded_depth_max_cm;
        L.ponded = cls.par.ponded_depth_max_cm;
      }
    }
    if (!(c.status & LGAR_FAIL_MASK))
      lg_move_begin(L, L.volin);
    else {
      L.nwf = 0;
      L.wf = 0;  // skip the front loop (and the corrections) of a failed column
      L.me_phase = 0;
    }
  }
}

// From the end of the front loop to the end of the sub-cycle (bmi_lgar.cxx:553-805).  Returns 1 when the
// forcing step is complete, 0 when another sub-cycle has been started, 2 when a correction search has
// to iterate first (call again afterwards).
LG_FN_NOINLINE int lg_cycle_end(const LgarConfig& cfg, LgLane& L) {
  LgCtx& c = L.c;
  LgFronts& f = L.f;
  LgScalars& s = L.s;
  const LgarColumnClass& cls = *L.cls;
  const double dt = L.pro.subtimestep_h;
  LgCycle& cy = L.cy;
  if ((L.nwf != 0 || L.create) && L.me_phase != 3) {
    if (lg_move_end(L)) return 2;
    L.temp_rch = L.bottom_flux;
  }
  if (L.create) {  // :565-576, then the bookkeeping branch :619-632
    double dry_depth = lg_calc_dry_depth(c, dt);
    lg_create_surficial_front(c, &L.ponded, &L.volin, dry_depth, f.theta[0]);
    if (L.ponded < cls.par.ponded_depth_max_cm) {
      L.volon_sub = L.ponded;
      L.ponded = 0.0;
      L.runoff = 0.0;
    } else {
      L.runoff = (L.ponded - cls.par.ponded_depth_max_cm);
      L.volon_sub = cls.par.ponded_depth_max_cm;
      L.ponded = cls.par.ponded_depth_max_cm;
    }
  }
  if (!(c.status & LGAR_FAIL_MASK)) {
    lg_clean_redundant(c);
    int new_front = (L.pro.switch_caching && L.create) ? 1 : -1;
    lg_dzdt_calc(c, L.ponded, dt, L.pro.switch_caching, s.cache_count, new_front);
  }
  if (L.pro.switch_caching) s.cache_count = 1;
  double volrech = L.temp_rch;
  if (!cfg.free_drainage_to_CR) {
    volrech += L.free_drainage;
  } else {
    cy.fd_for_CR += L.free_drainage + volrech;
    volrech = 0.0;
  }
  cy.AET = L.AET;
  cy.volin = L.volin;
  cy.runoff = L.runoff;
  cy.volon_sub = L.volon_sub;
  cy.volrech = volrech;
  cy.volon_ts = L.volon_ts;
  L.volend_sub = lg_mass_bal(c);
  double local_mb = lg_cycle_tail(cfg, s, cy, dt, L.volend_sub, L.v);
  L.volon_ts = L.volon_sub;
  L.last_AET = L.AET;
  L.last_PET_sub = cy.PET_sub;
  L.last_volrech = volrech;
  if (!(fabs(local_mb) <= cfg.mbal_tol)) c.status |= (isnan(local_mb) ? LGAR_FAIL_NAN : LGAR_FAIL_MBAL);
  if (!(f.depth[0] > 0.0)) c.status |= LGAR_FAIL_TOP_DEPTH;
  if (!(c.status & LGAR_FAIL_MASK) && L.cycle < L.pro.subcycles) {
    L.cycle++;
    lg_cycle_begin(cfg, L);
    return 0;
  }
  int tbc = 0;  // :824-834
  for (int i = 0; i < f.n; i++) tbc += f.tb[i] ? 1 : 0;
  if (tbc > cls.num_layers) c.status |= LGAR_FAIL_TO_BOTTOM;
  s.volon = L.volon_ts;
  s.prev_aet = L.last_AET;
  s.prev_pet = L.last_PET_sub;
  s.prev_rech = L.last_volrech;
  s.storage = L.volend_sub;
  s.fd_dzdt = f.dzdt[lg_wf_free_drainage(f) - 1];
  s.status |= c.status;
  return 1;
}

// ---- stages -----------------------------------------------------------------------------------------
// HEAD: the column `col` needs an active step at hour t.  Ends in FRONT, or in FETCH when the step
// turned out to replay cached fluxes (force_all_active testing only).
LG_FN_NOINLINE void lg_stage_head(const LgarConfig& cfg, const LgarDeviceState& d, const LgarWindow& w, LgLane& L, unsigned long long* front_sum) {
  const int64_t col = L.col;
  const int t = L.t;
  double P = w.precip[(int64_t)t * w.series_stride + col], E = w.pet[(int64_t)t * w.series_stride + col];
  L.pro = lg_prologue(cfg, L.s, P);
  if (cfg.adaptive_timestep) L.s.timestep_h = L.pro.subtimestep_h;
  L.s.cache_count = L.pro.cache_count;
  adjust_forcing(cfg, P, E, L.s.status);
  L.P = P;
  L.E = E;
  zero_vols(L.v);
  if (L.pro.cache_fluxes) {
    L.s.cache_fluxes = 1;
    if (w.paths) { LgCtx pc; pc.paths = w.paths; LG_PATH(pc, LGAR_PATH_CACHED_STEP); }
    lg_step_cached(cfg, L.s, L.pro, P, E, L.v);
    if (!(fabs(L.v.local_mb) <= cfg.mbal_tol)) L.s.status |= (isnan(L.v.local_mb) ? LGAR_FAIL_NAN : LGAR_FAIL_MBAL);
    lg_epi_step(cfg, d, w, t, col, L.ep, L.v, L.s.status, t == w.K - 1);
    if (w.nf_sink) w.nf_sink[(int64_t)t * w.sink_stride + col] = L.f.n;
    if (front_sum) *front_sum += (unsigned)L.f.n;
    L.t = t + 1;
    L.stage = LG_ST_FETCH;
    return;
  }
  L.s.cache_fluxes = 0;
  L.c.cfg = &cfg;
  L.c.cls = L.cls;
  L.c.f = &L.f;
  L.c.ff = L.ffv;
  L.c.paths = w.paths;
  L.c.use_ff = w.use_ff != 0;
  L.c.status = 0;
  L.volon_ts = L.s.volon;
  L.volend_sub = L.s.storage;
  L.last_AET = L.last_PET_sub = L.last_volrech = 0.0;
  L.cycle = 1;
  lg_cycle_begin(cfg, L);
  L.stage = LG_ST_FRONT;
}

// FRONT: front-loop iterations; ends in SEARCH or TAIL
LG_FN void lg_stage_front(LgLane& L) {
  L.stage = lg_move_fronts(L) ? LG_ST_SEARCH : LG_ST_TAIL;
  // The front loop is over: what kind of correction the list needs first (lgar.cxx:1777-1790) is known here -- lg_move_end would
  // find the same as its first act -- and the TAIL queue is binned by it (slots that need a merge / crossing / dry-over-wet fix
  // take long branches of their own).
  if (L.stage == LG_ST_TAIL && (L.nwf != 0 || L.create) && L.me_phase == 0) {
    L.me_correction = lg_correction_type(L.c);
    L.me_phase = 1;
  }
}

// TAIL: ends in FRONT (next sub-cycle), CORR, or FETCH (step finished: GIUH, cumulative volumes, outputs)
LG_FN_NOINLINE void lg_stage_tail(const LgarConfig& cfg, const LgarDeviceState& d, const LgarWindow& w, LgLane& L, unsigned long long* front_sum) {
  const int r = lg_cycle_end(cfg, L);
  if (r == 0) {
    L.stage = LG_ST_FRONT;
    return;
  }
  if (r == 2) {
    L.stage = LG_ST_CORR;
    return;
  }
  const int64_t col = L.col;
  const int t = L.t;
  lg_epi_step(cfg, d, w, t, col, L.ep, L.v, L.s.status, t == w.K - 1);
  if (w.nf_sink) w.nf_sink[(int64_t)t * w.sink_stride + col] = L.f.n;
  if (front_sum) *front_sum += (unsigned)L.f.n;
  L.t = t + 1;
  L.stage = LG_ST_FETCH;
}

// FETCH: walk the resident column through the hours that need no flux computation; when it has finished the
// window write its state back and (unless the slot is bound to one column) take the next column.  Ends in
// HEAD (active step at hour L.t) or DONE.  L may live in HBM: the hot loop works on register copies.
LG_FN_NOINLINE void lg_stage_fetch(const LgarConfig& cfg, const LgarDeviceState& d, const LgarWindow& w, LgLane& L, unsigned long long* front_sum) {
  for (;;) {
    if (L.col >= 0 && L.t >= w.K) {  // this column has finished the window
      lg_lane_export(cfg, d, L.col, L);
#ifndef LGAR_HOSTSIM
      if (w.done_columns) atomicAdd(w.done_columns, 1);
#endif
      if (w.work_trace) {
        w.work_trace[L.col] = L.tr_iters;
        w.work_trace[w.trace_stride + L.col] = L.tr_kcyc;
      }
      L.col = -1;
    }
    if (L.col < 0) {
      if (w.slot_is_column) {
        L.stage = LG_ST_DONE;
        return;
      }
#ifdef LGAR_HOSTSIM
      const int64_t idx = (*w.tile_counter)++;
#else
      const int64_t idx = atomicAdd(w.tile_counter, 1);
#endif
      if (idx >= (w.col_list ? (int64_t)*w.col_count : d.ncol)) {
        L.stage = LG_ST_DONE;
        return;
      }
      const int64_t col = w.col_list ? (int64_t)w.col_list[idx] : idx;
      L.col = col;
      L.t = 0;
      L.tr_iters = 0;
      L.tr_kcyc = 0;
      lg_lane_import(cfg, d, col, L);
    }
    LgScalars s = L.s;
    LgEpi ep = L.ep;
    const int t1 = lg_advance_resident(cfg, d, w, L.cls, L.col, L.t, L.f.n, s, ep, front_sum);
    L.s = s;
    L.ep = ep;
    L.t = t1;
    if (t1 < lg_hours_present(w)) {
      L.stage = LG_ST_HEAD;
      return;
    }
    if (t1 < w.K) {  // the forcing of hour t1 has not arrived yet: wait in the FETCH queue
#ifndef LGAR_HOSTSIM
      if (w.waiting) atomicAdd(w.waiting, 1);
#endif
      L.stage = LG_ST_FETCH;
      return;
    }
  }
}

#endif  // LGAR_LANES_CUH
