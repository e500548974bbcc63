/* oracle/lgar_oracle.c -- TEST INFRASTRUCTURE ONLY (see lgar_oracle.h).
 *
 * Plain-C restatement of the LGAR-C Update() hot path.  Each function cites the reference
 * file:line it follows.  Floating-point expressions keep the reference's association order
 * (compile with -ffp-contract=off) so that the oracle tracks a -O2 build of the reference to
 * the last bit on glibc.  Parity is pinned by tests/test_oracle_vs_ref.py (step-by-step against
 * oracle/_ref) and tests/test_oracle_golden.py (committed fixtures generated from oracle/_ref).
 */
#include "lgar_oracle.h"

#include <math.h>
#include <stdio.h>
#include <string.h>

/* reference src/lgar.cxx:63-70 */
#define THRESHOLD_NO_MOISTURE_DIFF 1.E-15
#define MBAL_ITERATIVE_TOLERANCE 1.E-10
#define MAX_ITER_MBAL_LOOP 1.E5
#define MAX_ITER_SATURATION_MBAL_LOOP 1.E4
#define TRUNCATION_DEPTH 1.E-9
#define FACTOR_LIMITS_LAYER_CROSSING_SPEED 2.0
#define DEPTH_AVOIDS_SAME_WF_DEPTH 1.E-6
#define PSI_UPPER_LIM 1.E7
/* reference src/bmi_lgar.cxx:22-57 */
#define PRECIP_THRESHOLD_MM_PER_H 1.0e-6
#define PET_EPSILON 1.0e-8
#define AET_PET_RATIO_THRESHOLD 0.75
#define VOLON_TIMESTEP_THRESHOLD_CM 1.0e-16
#define BOTTOM_BDY_FLUX_THRESHOLD_CM 1.0e-4
#define THRESHOLD_DZDT_CM_PER_H 1.0e0
#define NUM_TIMESTEPS_BEFORE_RESET_CACHE 24
#define SMALL_EPS 1.E-12

typedef struct {
  const lgo_params* p;
  lgo_state* s;
} ctx_t;

/* ------------------------------------------------------------------------------------------ */
/* L1 constitutive relations: reference src/soil_funcs.cxx                                     */
/* ------------------------------------------------------------------------------------------ */

/* soil_funcs.cxx:147-150 */
double lgo_theta_from_h(double h, const lgo_soil* s) {
  return (1.0 / (pow(1.0 + pow(s->alpha * h, s->n), s->m)) * (s->theta_e - s->theta_r) + s->theta_r);
}
/* soil_funcs.cxx:155-159 (+ util_funcs.cxx:27) */
double lgo_Se_from_h(double h, double alpha, double m, double n) {
  if (fabs(h) < 1.0E-10) return 1.0;
  return (1.0 / (pow(1.0 + pow(alpha * h, n), m)));
}
/* soil_funcs.cxx:164-167 */
double lgo_K_from_Se(double Se, double Ks, double m) {
  return (Ks * sqrt(Se) * pow(1.0 - pow(1.0 - pow(Se, 1.0 / m), m), 2.0));
}
/* soil_funcs.cxx:172-179 */
double lgo_h_from_Se(double Se, double alpha, double m, double n) {
  double result = 1.0 / alpha * pow(pow(Se, -1.0 / m) - 1.0, 1.0 / n);
  if (result > 1.E20) result = 1.E20;
  return result;
}
/* soil_funcs.cxx:184-187 */
static double Se_from_theta(double theta, double e, double r) { return ((theta - r) / (e - r)); }

static double theta_h(ctx_t* c, double h, const lgo_soil* s) {
  c->s->n_theta_calls++;
  return lgo_theta_from_h(h, s);
}

/* soil_funcs.cxx:34-140.  NB closed form: Se_f comes from theta1 and Se_i from theta2. */
double lgo_Geff(int use_closed_form_G, double theta1, double theta2, const lgo_soil* s, int nint) {
  double Geff;
  double theta_e = s->theta_e, theta_r = s->theta_r;
  double vg_alpha = s->alpha, vg_n = s->n, vg_m = s->m, Ksat = s->Ksat;
  if (!use_closed_form_G) {
    double h_i, h_f, Se_i, Se_f, h2, dh, Se1, Se2, K1, K2;
    Se_i = Se_from_theta(theta1, theta_e, theta_r);
    Se_f = Se_from_theta(theta2, theta_e, theta_r);
    h_i = lgo_h_from_Se(Se_i, vg_alpha, vg_m, vg_n);
    h_f = lgo_h_from_Se(Se_f, vg_alpha, vg_m, vg_n);
    dh = (h_i - h_f) / (double)nint;
    dh = dh * 0.01;
    Geff = 0.0;
    Se1 = Se_i;
    K1 = lgo_K_from_Se(Se1, Ksat, vg_m);
    h2 = h_f + dh;
    while (h2 < h_i) {
      double prior_h2 = h2;
      Se2 = lgo_Se_from_h(h2, vg_alpha, vg_m, vg_n);
      K2 = lgo_K_from_Se(Se2, Ksat, vg_m);
      if ((K1 - K2) / K2 > 0.02)
        dh = dh * 0.5;
      else
        dh = dh * 10.0;
      if (h2 < h_i) {
        Geff += (K1 + K2) * dh / 2.0;
      } else { /* unreachable in the reference too (h2 < h_i is the loop guard); kept for fidelity */
        dh = h_i - prior_h2;
        Se2 = lgo_Se_from_h(h_i, vg_alpha, vg_m, vg_n);
        K2 = lgo_K_from_Se(Se2, Ksat, vg_m);
        Geff += (K1 + K2) * dh / 2.0;
      }
      K1 = K2;
      h2 += dh;
    }
    Geff = fabs(Geff / Ksat);
    return Geff;
  } else {
    double lambda = s->lambda;
    double Se_f = Se_from_theta(theta1, theta_e, theta_r);
    double Se_i = Se_from_theta(theta2, theta_e, theta_r);
    double H_c = s->psib * ((2 + 3 * lambda) / (1 + 3 * lambda));
    Geff = H_c * (pow(Se_i, (3 + 1 / lambda)) - pow(Se_f, (3 + 1 / lambda))) /
           (1 - pow(Se_f, (3 + 1 / lambda)));
    if (isinf(Geff)) Geff = H_c;
    if (isnan(Geff)) Geff = H_c;
    return Geff;
  }
}

/* ------------------------------------------------------------------------------------------ */
/* soil table: reference src/lgar.cxx:2674-2758                                                */
/* ------------------------------------------------------------------------------------------ */
void lgo_soil_derive(lgo_soil* s, double theta_r, double theta_e, double alpha, double n,
                     double Ksat, double wilting_point_psi_cm, int log_mode) {
  double m, p, lambda;
  if (log_mode) { /* lgar.cxx:2714-2717 */
    alpha = pow(10.0, alpha);
    Ksat = pow(10.0, Ksat);
  }
  s->theta_r = theta_r;
  s->theta_e = theta_e;
  s->alpha = alpha;
  s->n = n;
  s->m = 1 - 1 / n; /* lgar.cxx:2718 */
  s->Ksat = Ksat;
  s->theta_wp = lgo_theta_from_h(wilting_point_psi_cm, s);
  if (1.0 < n) { /* lgar.cxx:2729-2736 */
    m = 1.0 - 1.0 / n;
    p = 1.0 + 2.0 / m;
    s->lambda = 2.0 / (p - 3.0);
    s->psib = (p + 3.0) * (147.8 + 8.1 * p + 0.092 * p * p) /
              (2.0 * alpha * p * (p - 1.0) * (55.6 + 7.4 * p + p * p));
  }
  lambda = s->lambda;
  s->h_min = s->psib * (2.0 + 3.0 / lambda) / (1.0 + 3.0 / lambda); /* lgar.cxx:2746 */
}

/* Quirk Q2: sscanf("%s %lf...") stops the name at the first blank, so a row whose quoted name
 * contains a space fails to convert any number and silently inherits the previous row's values. */
int lgo_read_soil_file(const char* path, int num_soil_types, double wilting_point_psi_cm,
                       int log_mode, lgo_soil* table) {
  FILE* fp = fopen(path, "r");
  char line[256], name[256];
  double theta_e = 0, theta_r = 0, vg_n = 0, alpha = 0, Ksat = 0;
  int nread = 0, soil = 1;
  if (!fp) return -1;
  if (!fgets(line, 255, fp)) {
    fclose(fp);
    return 0;
  }
  while (fgets(line, 255, fp) != NULL) {
    sscanf(line, "%s %lf %lf %lf %lf %lf", name, &theta_r, &theta_e, &alpha, &vg_n, &Ksat);
    memset(&table[soil], 0, sizeof(lgo_soil));
    if (soil > 1) { /* the reference's struct keeps stale lambda/psib if n <= 1; mirror that */
      table[soil].lambda = table[soil - 1].lambda;
      table[soil].psib = table[soil - 1].psib;
    }
    lgo_soil_derive(&table[soil], theta_r, theta_e, alpha, vg_n, Ksat, wilting_point_psi_cm, log_mode);
    nread++;
    soil++;
    if (num_soil_types == nread) break;
  }
  fclose(fp);
  return nread;
}

/* ------------------------------------------------------------------------------------------ */
/* L0 list behaviour on a fixed array: reference src/linked_list.cxx                           */
/* ------------------------------------------------------------------------------------------ */

/* listDeleteFront, linked_list.cxx:208-293.  idx is 0-based.  Returns the index of the front that
 * followed the deleted one (== idx).  The psi/theta/K re-propagation through to_bottom fronts
 * (:267-288, un-frozen Ksat, quirk Q11) is skipped when the head is deleted. */
static int front_delete(ctx_t* c, int idx) {
  lgo_state* s = c->s;
  int i;
  if (idx < 0 || idx >= s->n) {
    s->status |= LGO_FAIL_LIST;
    return idx;
  }
  if (idx == s->n - 1) s->status |= LGO_FAIL_LIST; /* reference dereferences NULL here */
  for (i = idx; i < s->n - 1; i++) s->f[i] = s->f[i + 1];
  s->n--;
  if (idx != 0) {
    int wf;
    for (wf = s->n - 1; wf != 0; wf--) {
      lgo_front* cur = &s->f[wf - 1];
      lgo_front* nxt = &s->f[wf];
      if (cur->to_bottom) {
        const lgo_soil* sp = &c->p->soil[cur->layer];
        double Se;
        cur->psi = nxt->psi;
        cur->theta = theta_h(c, cur->psi, sp);
        Se = Se_from_theta(cur->theta, sp->theta_e, sp->theta_r);
        cur->K = lgo_K_from_Se(Se, sp->Ksat, sp->m);
      }
    }
  }
  return idx;
}

/* listInsertFirst, linked_list.cxx:97-123 (psi and K are left for the caller to set) */
static void front_insert_first(ctx_t* c, double depth, double theta, int layer, int to_bottom) {
  lgo_state* s = c->s;
  int i;
  if (s->n >= LGO_MAX_FRONTS) {
    s->status |= LGO_FAIL_OVERFLOW;
    return;
  }
  for (i = s->n; i > 0; i--) s->f[i] = s->f[i - 1];
  s->n++;
  s->f[0].depth = depth;
  s->f[0].theta = theta;
  s->f[0].layer = layer;
  s->f[0].to_bottom = to_bottom;
  s->f[0].dzdt = 0.0;
  s->f[0].psi = 0.0;
  s->f[0].K = 0.0;
}

/* ------------------------------------------------------------------------------------------ */
/* L2 physics: reference src/lgar.cxx                                                          */
/* ------------------------------------------------------------------------------------------ */

/* lgar_calc_mass_bal, lgar.cxx:2632-2668 */
double lgo_calc_mass_bal(const double* cum, const lgo_front* f, int n) {
  double sum = 0.0;
  int i;
  for (i = 0; i < n; i++) {
    double base_depth = cum[f[i].layer - 1];
    if (i + 1 < n) {
      if (f[i + 1].layer == f[i].layer)
        sum += (f[i].depth - base_depth) * (f[i].theta - f[i + 1].theta);
      else
        sum += (f[i].depth - base_depth) * f[i].theta;
    } else {
      sum += f[i].theta * (f[i].depth - base_depth);
    }
  }
  return sum;
}
static double mass_bal(ctx_t* c) { return lgo_calc_mass_bal(c->p->cum_thickness, c->s->f, c->s->n); }

/* wetting_front_free_drainage, lgar.cxx:1227-1254 (returns a 1-based front number) */
static int wf_free_drainage(ctx_t* c) {
  const lgo_state* s = c->s;
  int wf = 1, i;
  for (i = 0; i < s->n; i++) {
    if (i + 1 < s->n) {
      if (s->f[i].layer == s->f[i + 1].layer)
        break;
      else
        wf++;
    }
  }
  if (wf > s->n) wf--;
  return wf;
}

/* calc_min_water_possible_for_free_drainage_wetting_front, lgar.cxx:3197-3232 */
static double min_water_fd(ctx_t* c, int wf_fd) {
  const lgo_state* s = c->s;
  double min_storage = 0.0, previous_depth = 0.0;
  int i;
  for (i = 0; i < s->n; i++) {
    const lgo_soil* sp = &c->p->soil[s->f[i].layer];
    double min_theta = theta_h(c, PSI_UPPER_LIM, sp);
    min_storage += min_theta * (s->f[i].depth - previous_depth);
    if (i + 1 == wf_fd) break;
    previous_depth = s->f[i].depth;
  }
  return min_storage;
}
/* calc_storage_in_free_drainage_wetting_front, lgar.cxx:3234-3252 */
static double storage_fd(ctx_t* c, int wf_fd) {
  const lgo_state* s = c->s;
  double storage = 0.0, previous_depth = 0.0;
  int i;
  for (i = 0; i < s->n; i++) {
    storage += s->f[i].theta * (s->f[i].depth - previous_depth);
    if (i + 1 == wf_fd) break;
    previous_depth = s->f[i].depth;
  }
  return storage;
}

/* lgar_theta_mass_balance, lgar.cxx:2952-3112.  delta_theta / delta_thickness are 1-based. */
static double theta_mass_balance(ctx_t* c, int layer_num, double psi_cm, double new_mass,
                                 double prior_mass, double precip_mass_to_add, double* AET_demand_cm,
                                 const double* delta_theta, const double* delta_thickness) {
  const lgo_params* p = c->p;
  const lgo_soil* sp = &p->soil[layer_num];
  double psi_cm_loc = psi_cm;
  double delta_mass = fabs(new_mass - prior_mass);
  double factor = fmax(1.0, psi_cm / 10.0);
  int switched = 0;
  double theta = 0;
  double psi_cm_loc_prev = psi_cm_loc;
  double delta_mass_prev = delta_mass;
  int count_no_mass_change = 0;
  int break_no_mass_change = 5;
  int wanted_to_saturate_flag = 0;
  int iter = 0;
  int iter_aug_flag = 0, iter_aug_flag_extreme = 0;
  int first_speedup_thresh = (int)(MAX_ITER_MBAL_LOOP / 100);
  int second_speedup_thresh = (int)(MAX_ITER_MBAL_LOOP / 10);
  if (psi_cm > 1.E4) factor = psi_cm;

  if (delta_mass <= MBAL_ITERATIVE_TOLERANCE) {
    theta = theta_h(c, psi_cm_loc, sp);
    return theta;
  }

  while (delta_mass > MBAL_ITERATIVE_TOLERANCE) {
    double theta_layer, mass_layers = 0.0;
    int k;
    iter++;
    c->s->n_tmb_iters++;
    if (iter > first_speedup_thresh && !iter_aug_flag) {
      factor = factor * 100;
      iter_aug_flag = 1;
    }
    if (iter > second_speedup_thresh && !iter_aug_flag_extreme) {
      factor = factor * 100;
      iter_aug_flag_extreme = 1;
    }
    if (new_mass > prior_mass) {
      psi_cm_loc += 0.1 * factor;
      switched = 0;
    } else {
      if (!switched) {
        switched = 1;
        factor = factor * 0.1;
      }
      psi_cm_loc_prev = psi_cm_loc;
      psi_cm_loc -= 0.1 * factor;
      if (psi_cm_loc < 0.0) wanted_to_saturate_flag = 1;
      if (psi_cm_loc < 0 && psi_cm_loc_prev != 0) psi_cm_loc = psi_cm_loc_prev * 0.1;
    }
    if (psi_cm_loc < 0.0) {
      psi_cm_loc = 0.0;
      factor = factor * 0.5;
    }
    theta = theta_h(c, psi_cm_loc, sp);
    mass_layers += delta_thickness[layer_num] * (theta - delta_theta[layer_num]);
    for (k = 1; k < layer_num; k++) {
      theta_layer = theta_h(c, psi_cm_loc, &p->soil[k]);
      mass_layers += delta_thickness[k] * (theta_layer - delta_theta[k]);
    }
    new_mass = mass_layers;
    delta_mass = fabs(new_mass - prior_mass);

    if (fabs(psi_cm_loc - psi_cm_loc_prev) < 1E-15 && factor < 1E-13) break;
    if (fabs(delta_mass - delta_mass_prev) < 1E-15)
      count_no_mass_change++;
    else
      count_no_mass_change = 0;
    if (count_no_mass_change == break_no_mass_change && precip_mass_to_add < 1.E-12) break;
    if (iter > MAX_ITER_MBAL_LOOP) break;
    if ((psi_cm_loc > PSI_UPPER_LIM) && (iter > first_speedup_thresh)) break;
    if (psi_cm_loc <= 0 && psi_cm_loc_prev < 0) break;
    delta_mass_prev = delta_mass;
  }

  { /* which exit ended the loop (instrumentation; the conditions are those of the loop, in its order) */
    lgo_state* st = c->s;
    st->n_path[LGO_PATH_SEARCH]++;
    if (!(delta_mass > MBAL_ITERATIVE_TOLERANCE)) st->n_path[LGO_PATH_SEARCH_CONVERGED]++;
    else if (fabs(psi_cm_loc - psi_cm_loc_prev) < 1E-15 && factor < 1E-13) st->n_path[LGO_PATH_SEARCH_STEP_EXIT]++;
    else if (count_no_mass_change == break_no_mass_change && precip_mass_to_add < 1.E-12) st->n_path[LGO_PATH_SEARCH_NOCHANGE]++;
    else if (iter > MAX_ITER_MBAL_LOOP) st->n_path[LGO_PATH_SEARCH_MAXITER]++;
    else if ((psi_cm_loc > PSI_UPPER_LIM) && (iter > first_speedup_thresh)) st->n_path[LGO_PATH_SEARCH_PSI_UPPER]++;
    else st->n_path[LGO_PATH_SEARCH_PSI_ZERO]++;
  }
  if ((delta_mass > MBAL_ITERATIVE_TOLERANCE) && (!wanted_to_saturate_flag))
    *AET_demand_cm = *AET_demand_cm - fabs(delta_mass - MBAL_ITERATIVE_TOLERANCE);
  if ((theta >= sp->theta_e) && (psi_cm_loc != 0.0)) *AET_demand_cm = 0.0;
  return theta;
}

/* lgar_theta_mass_balance_correction, lgar.cxx:3262-3435.  front_num is 1-based. */
static void theta_mass_balance_correction(ctx_t* c, int front_num, double prior_mass) {
  lgo_state* s = c->s;
  const lgo_params* p = c->p;
  int cur = front_num - 1; /* index of `current` */
  double new_mass, psi_cm, psi_cm_loc, delta_mass, factor, psi_cm_loc_prev, delta_mass_prev;
  int switched = 0, count_no_mass_change = 0, break_no_mass_change = 5;
  int iter = 0, iter_aug_flag = 0, iter_aug_flag_extreme = 0;
  int first_speedup_thresh = (int)(MAX_ITER_MBAL_LOOP / 100);
  int second_speedup_thresh = (int)(MAX_ITER_MBAL_LOOP / 10);

  if (cur < 0 || cur >= s->n) {
    s->status |= LGO_FAIL_LIST;
    return;
  }
  if (cur + 1 > 1) { /* :3266-3271 */
    if (s->f[cur - 1].to_bottom) cur = cur - 1;
  }
  while (s->f[cur].to_bottom) { /* :3272-3281 */
    if (cur + 1 == 1) break;
    cur = cur - 1;
    if (!s->f[cur].to_bottom) {
      cur = cur + 1;
      break;
    }
  }
  new_mass = mass_bal(c);
  psi_cm = s->f[cur].psi;
  psi_cm_loc = s->f[cur].psi;
  delta_mass = fabs(new_mass - prior_mass);
  factor = fmax(1.0, psi_cm / 10.0);
  if (psi_cm > 1.E4) factor = psi_cm;
  psi_cm_loc_prev = psi_cm_loc;
  delta_mass_prev = delta_mass;

  if (delta_mass > MBAL_ITERATIVE_TOLERANCE) c->s->n_path[LGO_PATH_CORR_SEARCH]++;
  while (delta_mass > MBAL_ITERATIVE_TOLERANCE) {
    int nx, before;
    iter++;
    c->s->n_tmb_iters++;
    if (iter > first_speedup_thresh && !iter_aug_flag) {
      factor = factor * 100;
      iter_aug_flag = 1;
    }
    if (iter > second_speedup_thresh && !iter_aug_flag_extreme) {
      factor = factor * 100;
      iter_aug_flag_extreme = 1;
    }
    if (new_mass > prior_mass) {
      psi_cm_loc += 0.1 * factor;
      switched = 0;
    } else {
      if (!switched) {
        switched = 1;
        factor = factor * 0.1;
      }
      psi_cm_loc_prev = psi_cm_loc;
      psi_cm_loc -= 0.1 * factor;
      if (psi_cm_loc < 0 && psi_cm_loc_prev != 0) psi_cm_loc = psi_cm_loc_prev * 0.1;
    }
    if (psi_cm_loc < 0.0) {
      psi_cm_loc = 0.0;
      factor = factor * 0.5;
    }
    s->f[cur].psi = psi_cm_loc;
    s->f[cur].theta = theta_h(c, psi_cm_loc, &p->soil[s->f[cur].layer]);

    nx = cur + 1; /* :3356-3397 */
    before = cur;
    if (nx < s->n) {
      while (s->f[nx].to_bottom) {
        s->f[nx].psi = s->f[before].psi;
        s->f[nx].theta = theta_h(c, s->f[nx].psi, &p->soil[s->f[nx].layer]);
        before = nx;
        nx = nx + 1;
        if (nx >= s->n) break;
      }
      if (nx < s->n) {
        if (!s->f[nx].to_bottom && s->f[before].to_bottom) {
          s->f[nx].psi = s->f[before].psi;
          s->f[nx].theta = theta_h(c, s->f[nx].psi, &p->soil[s->f[nx].layer]);
        }
      }
    }
    new_mass = mass_bal(c);
    delta_mass = fabs(new_mass - prior_mass);

    if (fabs(psi_cm_loc - psi_cm_loc_prev) < 1E-15 && factor < 1E-13) break;
    if (fabs(delta_mass - delta_mass_prev) < 1E-15)
      count_no_mass_change++;
    else
      count_no_mass_change = 0;
    if (count_no_mass_change == break_no_mass_change) break;
    if (iter > MAX_ITER_MBAL_LOOP) break;
    if (psi_cm_loc <= 0 && psi_cm_loc_prev < 0) break;
    if ((psi_cm_loc > PSI_UPPER_LIM) && (iter > first_speedup_thresh)) break;
    delta_mass_prev = delta_mass;
  }
}

/* lgar_merge_wetting_fronts, lgar.cxx:1875-1953 */
static void merge_wetting_fronts(ctx_t* c) {
  lgo_state* s = c->s;
  int cur = 0, wf;
  for (wf = 1; wf != s->n; wf++) {
    int nx = cur + 1, nn = cur + 2;
    if (nx >= s->n) {
      s->status |= LGO_FAIL_LIST;
      return;
    }
    if ((s->f[cur].depth > s->f[nx].depth) && (s->f[cur].layer == s->f[nx].layer) && !s->f[nx].to_bottom) {
      const lgo_soil* sp;
      double current_mass_this_layer, Se;
      if (nn >= s->n) {
        s->status |= LGO_FAIL_LIST;
        return;
      }
      current_mass_this_layer = s->f[cur].depth * (s->f[cur].theta - s->f[nx].theta) +
                                s->f[nx].depth * (s->f[nx].theta - s->f[nn].theta);
      s->f[cur].depth = current_mass_this_layer / (s->f[cur].theta - s->f[nn].theta);
      if (!(s->f[cur].depth > 0.0)) s->status |= LGO_FAIL_MERGE_DEPTH;
      sp = &c->p->soil[s->f[cur].layer];
      Se = Se_from_theta(s->f[cur].theta, sp->theta_e, sp->theta_r);
      s->f[cur].psi = lgo_h_from_Se(Se, sp->alpha, sp->m, sp->n);
      s->f[cur].K = lgo_K_from_Se(Se, sp->Ksat, sp->m);
      front_delete(c, nx);
    }
    cur = cur + 1;
    if (wf + 1 != s->n && cur >= s->n) { /* the reference would walk off the list */
      s->status |= LGO_FAIL_LIST;
      return;
    }
  }
}

/* lgar_wetting_fronts_cross_layer_boundary, lgar.cxx:1963-2167 */
static void cross_layer_boundary(ctx_t* c) {
  lgo_state* s = c->s;
  const lgo_params* p = c->p;
  const double* cum = p->cum_thickness;
  int num_layers = p->num_layers;
  int cur = 0, wf;
  int cross_necessary = 0, theta_correction_necessary = 0, front_for_cross = -1;

  for (wf = 1; wf != s->n; wf++) {
    int layer_num = s->f[cur].layer;
    const lgo_soil* sp = &p->soil[layer_num];
    int nx = cur + 1, nn = cur + 2;
    if (nx >= s->n) {
      s->status |= LGO_FAIL_LIST;
      return;
    }
    if (s->f[cur].depth > cum[layer_num] && (s->f[nx].depth == cum[layer_num]) && (s->f[nx].to_bottom) &&
        (layer_num != num_layers)) {
      double current_theta = fmin(sp->theta_e, s->f[cur].theta);
      double overshot_depth = s->f[cur].depth - s->f[nx].depth;
      const lgo_soil* spn = &p->soil[layer_num + 1];
      double Se, theta_new, mbal_correction, mbal_Z_correction, depth_new;
      if (nn >= s->n) {
        s->status |= LGO_FAIL_LIST;
        return;
      }
      Se = Se_from_theta(s->f[cur].theta, sp->theta_e, sp->theta_r);
      s->f[cur].psi = lgo_h_from_Se(Se, sp->alpha, sp->m, sp->n);
      s->f[cur].K = lgo_K_from_Se(Se, sp->Ksat, sp->m);
      theta_new = theta_h(c, s->f[cur].psi, spn);
      mbal_correction = overshot_depth * (current_theta - s->f[nx].theta);
      mbal_Z_correction = mbal_correction / (theta_new - s->f[nn].theta);
      if (isinf(mbal_Z_correction)) mbal_Z_correction = cum[layer_num + 1] / (1e3);
      depth_new = cum[layer_num] + mbal_Z_correction;
      s->f[cur].depth = cum[layer_num];
      s->f[nx].theta = theta_new;
      s->f[nx].psi = s->f[cur].psi;
      s->f[nx].depth = depth_new;
      s->f[nx].layer = layer_num + 1;
      s->f[nx].dzdt = s->f[cur].dzdt;
      s->f[cur].dzdt = 0;
      s->f[cur].to_bottom = 1;
      s->f[nx].to_bottom = 0;
      cross_necessary = 1;
      front_for_cross = nx + 1; /* 1-based front number */
      if (depth_new > cum[num_layers] && s->f[nx].layer == num_layers) theta_correction_necessary = 1;
    }
    cur = cur + 1;
  }

  if (cross_necessary) {
    for (wf = s->n - 1; wf != 0; wf--) { /* :2058-2084, un-frozen Ksat (quirk Q11) */
      lgo_front* ct = &s->f[wf - 1];
      lgo_front* nt = &s->f[wf];
      const lgo_soil* sk = &p->soil[ct->layer];
      double Se;
      if (ct->to_bottom) {
        ct->psi = nt->psi;
        ct->theta = theta_h(c, ct->psi, sk);
        Se = Se_from_theta(ct->theta, sk->theta_e, sk->theta_r);
        ct->K = lgo_K_from_Se(Se, sk->Ksat, sk->m);
      } else {
        Se = Se_from_theta(ct->theta, sk->theta_e, sk->theta_r);
        ct->K = lgo_K_from_Se(Se, sk->Ksat, sk->m);
      }
    }
    if (theta_correction_necessary) { /* :2090-2159 */
      for (wf = s->n - 1; wf != 0; wf--) {
        lgo_front* ct;
        double prior_mass;
        if (wf - 1 >= s->n) {
          s->status |= LGO_FAIL_LIST;
          return;
        }
        ct = &s->f[wf - 1];
        prior_mass = mass_bal(c);
        if (ct->depth > cum[num_layers] && ct->layer == num_layers && wf == front_for_cross) {
          double current_mass;
          s->n_path[LGO_PATH_CROSS_LAYER_BOTTOM]++;
          ct->depth = cum[num_layers] + 1.E-6;
          theta_mass_balance_correction(c, wf, prior_mass);
          current_mass = mass_bal(c);
          if (prior_mass > (current_mass + 100. * MBAL_ITERATIVE_TOLERANCE)) {
            double depth_increment = 0.001;
            double max_increment = 1. / TRUNCATION_DEPTH;
            double max_depth = cum[num_layers] + 1. / TRUNCATION_DEPTH;
            double increment = depth_increment;
            int found_upper_bound = 0;
            int iter_one_direction = 0;
            while (prior_mass > current_mass && increment < max_increment && ct->depth < max_depth) {
              iter_one_direction++;
              if (iter_one_direction > MAX_ITER_MBAL_LOOP) break;
              ct->depth += increment;
              current_mass = mass_bal(c);
              if (current_mass >= prior_mass) {
                found_upper_bound = 1;
                break;
              }
              increment *= 2.0;
            }
            if (found_upper_bound) {
              double low = ct->depth - increment;
              double high = ct->depth;
              double mid;
              double current_mass2 = mass_bal(c);
              while (fabs(current_mass2 - prior_mass) > MBAL_ITERATIVE_TOLERANCE) {
                iter_one_direction++;
                mid = 0.5 * (low + high);
                ct->depth = mid;
                current_mass2 = mass_bal(c);
                if (current_mass2 >= prior_mass)
                  high = mid;
                else
                  low = mid;
                if (iter_one_direction > MAX_ITER_MBAL_LOOP) break;
              }
            }
          }
        }
      }
    }
  }
}

/* lgar_wetting_front_cross_domain_boundary, lgar.cxx:2176-2251 */
static double cross_domain_boundary(ctx_t* c) {
  lgo_state* s = c->s;
  const lgo_params* p = c->p;
  double domain_depth_cm = p->cum_thickness[p->num_layers];
  double bottom_flux_cm = 0.0;
  int length = s->n, cur = 0, wf;
  for (wf = 1; wf != length; wf++) {
    double bottom_flux_cm_temp;
    int nx, nn;
    if (wf >= s->n) break;
    bottom_flux_cm_temp = 0.0;
    nx = cur + 1;
    nn = cur + 2;
    if (nx >= s->n) {
      s->status |= LGO_FAIL_LIST;
      return bottom_flux_cm;
    }
    if (nn >= s->n && s->f[cur].depth >= domain_depth_cm) {
      const lgo_soil* sp = &p->soil[s->f[cur].layer];
      double Se_k;
      bottom_flux_cm_temp = (s->f[cur].theta - s->f[nx].theta) * (s->f[cur].depth - s->f[nx].depth);
      s->f[nx].theta = s->f[cur].theta;
      Se_k = Se_from_theta(s->f[cur].theta, sp->theta_e, sp->theta_r);
      s->f[nx].psi = lgo_h_from_Se(Se_k, sp->alpha, sp->m, sp->n);
      s->f[nx].K = lgo_K_from_Se(Se_k, sp->Ksat, sp->m);
      front_delete(c, cur);
      bottom_flux_cm += bottom_flux_cm_temp;
      break;
    }
    cur = nx;
    if (bottom_flux_cm_temp != 0.0) break;
  }
  return bottom_flux_cm;
}

/* lgar_check_dry_over_wet_wetting_fronts, lgar.cxx:2314-2338 */
static int check_dry_over_wet(ctx_t* c) {
  const lgo_state* s = c->s;
  int i;
  for (i = 0; i + 1 < s->n; i++)
    if ((s->f[i].theta <= s->f[i + 1].theta) && (s->f[i].layer == s->f[i + 1].layer)) return 1;
  return 0;
}

/* lgar_fix_dry_over_wet_wetting_fronts, lgar.cxx:2260-2307 */
static void fix_dry_over_wet(ctx_t* c, double* mass_change) {
  lgo_state* s = c->s;
  int cur = 0, nx = (s->n > 1) ? 1 : -1, l;
  for (l = 1; l <= s->n; l++) {
    if (nx >= 0) {
      if ((s->f[cur].theta <= s->f[nx].theta) && (s->f[cur].layer == s->f[nx].layer)) {
        double prior_mass = mass_bal(c), mass_after;
        cur = front_delete(c, cur);
        theta_mass_balance_correction(c, cur + 1, prior_mass);
        mass_after = mass_bal(c);
        *mass_change += (mass_after - prior_mass);
      }
      cur = cur + 1;
      if (cur >= s->n)
        nx = -1;
      else
        nx = (cur + 1 < s->n) ? cur + 1 : -1;
    }
  }
}

/* lgarto_correction_type_surf, lgar.cxx:3116-3169 */
static int correction_type_surf(ctx_t* c) {
  const lgo_state* s = c->s;
  const double* cum = c->p->cum_thickness;
  int num_layers = c->p->num_layers;
  int cur = 0, wf;
  for (wf = 1; wf != s->n; wf++) {
    int nx = cur + 1;
    int nn_is_null = (cur + 2 >= s->n);
    if (nx < s->n) {
      const lgo_front* a = &s->f[cur];
      const lgo_front* b = &s->f[nx];
      if ((a->theta > b->theta) && (a->depth > b->depth) && (a->layer == b->layer) && (!b->to_bottom)) return 1;
      if ((a->depth > cum[a->layer]) && (b->depth == cum[a->layer] && (b->to_bottom)) && (a->theta > b->theta) &&
          (a->layer != num_layers))
        return 2;
    }
    if (nn_is_null && (s->f[cur].depth > cum[s->f[cur].layer])) return 3;
    if (check_dry_over_wet(c)) return 4;
    cur = nx;
  }
  return 0;
}

/* lgar_clean_redundant_fronts, lgar.cxx:3171-3195 */
static void clean_redundant_fronts(ctx_t* c) {
  lgo_state* s = c->s;
  int cur = 0, wf;
  for (wf = 1; wf != s->n; wf++) {
    int nx = cur + 1;
    if ((s->f[cur].layer == s->f[nx].layer) && (fabs(s->f[cur].theta - s->f[nx].theta) < THRESHOLD_NO_MOISTURE_DIFF)) {
      s->n_path[LGO_PATH_CLEAN_REDUNDANT]++;
      front_delete(c, cur);
      break;
    }
    cur = nx;
  }
}

/* lgar_move_wetting_fronts, lgar.cxx:1269-1866.  old[] / n_old is state_previous. */
static double move_wetting_fronts(ctx_t* c, double timestep_h, double* free_drainage_subtimestep_cm,
                                  double* volin_cm, int wf_free_drainage_demand, double old_mass,
                                  double mass_correction_fd, double* AET_demand_cm, const lgo_front* old,
                                  int n_old) {
  lgo_state* s = c->s;
  const lgo_params* p = c->p;
  const double* cum = p->cum_thickness;
  int num_layers = p->num_layers;
  double column_depth = cum[num_layers];
  int number_of_wetting_fronts = s->n;
  int last_wetting_front_index = number_of_wetting_fronts;
  double precip_mass_to_add = (*volin_cm);
  double bottom_boundary_flux_cm = 0.0;
  double free_drainage_demand = *free_drainage_subtimestep_cm;
  double mass_change;
  int wf, correction;
  double delta_thetas[LGO_MAX_LAYERS + 2], delta_thickness[LGO_MAX_LAYERS + 2];

  *volin_cm = 0.0;

  for (wf = number_of_wetting_fronts; wf != 0; wf--) {
    int ci = wf - 1; /* index of `current` in the live list */
    int ni = wf;     /* index of `next` */
    int layer_num, layer_num_below;
    const lgo_soil* sp;
    double theta_e, theta_r, actual_ET_demand;
    if (ci >= s->n || ci >= n_old) {
      s->status |= LGO_FAIL_LIST;
      return bottom_boundary_flux_cm;
    }
    layer_num = s->f[ci].layer;
    sp = &p->soil[layer_num];
    theta_e = sp->theta_e;
    theta_r = sp->theta_r;
    if (wf == last_wetting_front_index)
      layer_num_below = layer_num + 1;
    else {
      if (ni >= s->n) {
        s->status |= LGO_FAIL_LIST;
        return bottom_boundary_flux_cm;
      }
      layer_num_below = s->f[ni].layer;
    }
    actual_ET_demand = *AET_demand_cm;

    /* deepest front of a layer, not the last front: psi continuity across the boundary :1379-1387 */
    if ((wf < last_wetting_front_index) && (layer_num_below != layer_num)) {
      s->f[ci].theta = theta_h(c, s->f[ni].psi, sp);
      s->f[ci].psi = s->f[ni].psi;
    }

    /* one front per layer: :1403-1480 */
    if (wf == number_of_wetting_fronts && layer_num_below != layer_num && number_of_wetting_fronts == num_layers) {
      double psi_cm_old, psi_cm, prior_mass, new_mass, theta_new, Se;
      int k;
      s->f[ci].depth += s->f[ci].dzdt * timestep_h;
      psi_cm_old = old[ci].psi;
      psi_cm = s->f[ci].psi;
      prior_mass = (old[ci].depth - cum[layer_num - 1]) * (old[ci].theta - 0.0);
      new_mass = (s->f[ci].depth - cum[layer_num - 1]) * (s->f[ci].theta - 0.0);
      for (k = 1; k < layer_num; k++) {
        const lgo_soil* sk = &p->soil[k];
        double theta_old = theta_h(c, psi_cm_old, sk);
        double theta_below_old = 0.0;
        double local_delta_theta_old = theta_old - theta_below_old;
        double layer_thickness = cum[k] - cum[k - 1];
        double theta, theta_below = 0.0;
        prior_mass += (layer_thickness * local_delta_theta_old);
        theta = theta_h(c, psi_cm, sk);
        new_mass += layer_thickness * (theta - theta_below);
        delta_thetas[k] = theta_below;
        delta_thickness[k] = layer_thickness;
      }
      delta_thetas[layer_num] = 0.0;
      delta_thickness[layer_num] = s->f[ci].depth - cum[layer_num - 1];
      if (wf_free_drainage_demand == wf)
        prior_mass += precip_mass_to_add - (free_drainage_demand + mass_correction_fd + actual_ET_demand);
      theta_new = theta_mass_balance(c, layer_num, psi_cm, new_mass, prior_mass, precip_mass_to_add,
                                     AET_demand_cm, delta_thetas, delta_thickness);
      actual_ET_demand = *AET_demand_cm;
      s->f[ci].theta = fmax(theta_r, fmin(theta_new, theta_e));
      Se = Se_from_theta(s->f[ci].theta, theta_e, theta_r);
      s->f[ci].psi = lgo_h_from_Se(Se, sp->alpha, sp->m, sp->n);
    }

    /* front inside a layer: :1489-1637 */
    if ((wf < last_wetting_front_index) && (layer_num == layer_num_below)) {
      double Se;
      if (ni >= n_old) {
        s->status |= LGO_FAIL_LIST;
        return bottom_boundary_flux_cm;
      }
      if (layer_num == 1) {
        double prior_mass = old[ci].depth * (old[ci].theta - old[ni].theta);
        if (wf_free_drainage_demand == wf)
          prior_mass += precip_mass_to_add - (free_drainage_demand + mass_correction_fd + actual_ET_demand);
        s->f[ci].depth += s->f[ci].dzdt * timestep_h;
        if (s->f[ci].depth > column_depth) s->f[ci].depth = column_depth + TRUNCATION_DEPTH;
        if (s->f[ci].dzdt == 0.0 && s->f[ci].to_bottom == 0) {
          /* a new front was just created, so don't update it */
        } else {
          if ((prior_mass / s->f[ci].depth + s->f[ni].theta) < theta_r) { /* :1519-1537 */
            double mass_before = mass_bal(c) - s->f[ci].depth * (s->f[ci].theta -
                                 (prior_mass / s->f[ci].depth + s->f[ni].theta));
            double mass_after;
            s->n_path[LGO_PATH_THETA_R_DELETE]++;
            front_delete(c, ci); /* `current` becomes the old `next`, which now sits at index ci */
            mass_after = mass_bal(c);
            *AET_demand_cm = *AET_demand_cm - fabs(mass_before - mass_after);
            actual_ET_demand = *AET_demand_cm;
          } else {
            s->f[ci].theta = fmax(theta_r, fmin(theta_e, prior_mass / s->f[ci].depth + s->f[ni].theta));
          }
        }
      } else {
        double psi_cm_old, psi_cm_below_old, psi_cm, psi_cm_below, prior_mass, new_mass, theta_new;
        int k;
        s->f[ci].depth += s->f[ci].dzdt * timestep_h;
        if (s->f[ci].depth > column_depth) s->f[ci].depth = column_depth + TRUNCATION_DEPTH;
        psi_cm_old = old[ci].psi;
        psi_cm_below_old = old[ni].psi;
        psi_cm = s->f[ci].psi;
        psi_cm_below = s->f[ni].psi;
        prior_mass = (old[ci].depth - cum[layer_num - 1]) * (old[ci].theta - old[ni].theta);
        new_mass = (s->f[ci].depth - cum[layer_num - 1]) * (s->f[ci].theta - s->f[ni].theta);
        for (k = 1; k < layer_num; k++) {
          const lgo_soil* sk = &p->soil[k];
          double theta_old = theta_h(c, psi_cm_old, sk);
          double theta_below_old = theta_h(c, psi_cm_below_old, sk);
          double local_delta_theta_old = theta_old - theta_below_old;
          double layer_thickness = (cum[k] - cum[k - 1]);
          double theta, theta_below;
          prior_mass += (layer_thickness * local_delta_theta_old);
          theta = theta_h(c, psi_cm, sk);
          theta_below = theta_h(c, psi_cm_below, sk);
          new_mass += layer_thickness * (theta - theta_below);
          delta_thetas[k] = theta_below;
          delta_thickness[k] = layer_thickness;
        }
        delta_thetas[layer_num] = s->f[ni].theta;
        delta_thickness[layer_num] = s->f[ci].depth - cum[layer_num - 1];
        if (wf_free_drainage_demand == wf)
          prior_mass += precip_mass_to_add - (free_drainage_demand + mass_correction_fd + actual_ET_demand);
        theta_new = theta_mass_balance(c, layer_num, psi_cm, new_mass, prior_mass, precip_mass_to_add,
                                       AET_demand_cm, delta_thetas, delta_thickness);
        actual_ET_demand = *AET_demand_cm;
        s->f[ci].theta = fmax(theta_r, fmin(theta_new, theta_e));
      }
      /* :1634-1635 -- after a deletion `current` is the former `next` (same layer, same soil) */
      Se = Se_from_theta(s->f[ci].theta, theta_e, theta_r);
      s->f[ci].psi = lgo_h_from_Se(Se, sp->alpha, sp->m, sp->n);
    }

    /* saturated infiltrating front: deepen it until the mass balance closes, :1644-1754 */
    if (wf == 1) {
      int fdi;
      double theta_e_k1, mass_timestep;
      wf_free_drainage_demand = wf_free_drainage(c);
      fdi = wf_free_drainage_demand - 1;
      theta_e_k1 = p->soil[s->f[fdi].layer].theta_e;
      mass_timestep = (old_mass + precip_mass_to_add) - (actual_ET_demand + free_drainage_demand + mass_correction_fd);
      if (!(old_mass > 0.0)) s->status |= LGO_FAIL_OLD_MASS;
      if (fabs(s->f[fdi].theta - theta_e_k1) < 1E-15) {
        double current_mass = mass_bal(c);
        double mass_balance_error = fabs(current_mass - mass_timestep);
        double factor = 1.0;
        int switched = 0;
        double depth_new;
        int iter = 0, iter_aug_flag = 0, break_flag = 0;
        int speedup_thresh = (int)(MAX_ITER_SATURATION_MBAL_LOOP / 10);
        if (fdi + 1 < s->n) {
          if (fabs(s->f[fdi].theta - s->f[fdi + 1].theta) < 0.01)
            factor = factor / fabs(s->f[fdi].theta - s->f[fdi + 1].theta);
        }
        depth_new = s->f[fdi].depth;
        if (fabs(mass_balance_error) > MBAL_ITERATIVE_TOLERANCE) s->n_path[LGO_PATH_SATURATED_DEPTH]++;
        while (fabs(mass_balance_error) > MBAL_ITERATIVE_TOLERANCE) {
          iter++;
          if (iter > MAX_ITER_SATURATION_MBAL_LOOP) {
            break_flag = 1;
            *AET_demand_cm = *AET_demand_cm + mass_balance_error;
            actual_ET_demand = *AET_demand_cm;
            break;
          }
          if ((iter > speedup_thresh) && (!iter_aug_flag)) {
            factor = factor * 100;
            iter_aug_flag = 1;
          }
          if (current_mass < mass_timestep) {
            depth_new += 0.01 * factor;
            switched = 0;
          } else {
            if (!switched) {
              switched = 1;
              factor = factor * 0.1;
            }
            depth_new -= 0.01 * factor;
          }
          if ((s->f[fdi].to_bottom == 1) && (s->f[fdi].layer == num_layers)) depth_new = cum[num_layers];
          s->f[fdi].depth = depth_new;
          current_mass = mass_bal(c);
          mass_balance_error = fabs(current_mass - mass_timestep);
        }
        if (depth_new < TRUNCATION_DEPTH) {
          depth_new = TRUNCATION_DEPTH;
          s->f[fdi].depth = depth_new;
        }
        if (isinf(s->f[fdi].depth)) {
          s->f[fdi].depth = cum[num_layers] + 1.E-6;
          break_flag = 1;
        }
        if (break_flag) {
          current_mass = mass_bal(c);
          mass_timestep = (old_mass + precip_mass_to_add) - (actual_ET_demand + free_drainage_demand + mass_correction_fd);
          mass_balance_error = mass_timestep - current_mass;
          bottom_boundary_flux_cm += mass_balance_error;
        }
      }
    }
  }

  /* merging / crossing corrections until none is left, :1777-1814 */
  mass_change = 0.0;
  correction = correction_type_surf(c);
  while (correction != 0) {
    if (s->status & (LGO_FAIL_LIST | LGO_FAIL_OVERFLOW)) break;
    s->n_path[correction == 1 ? LGO_PATH_MERGE : correction == 2 ? LGO_PATH_CROSS_LAYER : correction == 3 ? LGO_PATH_CROSS_BOTTOM : LGO_PATH_DRY_OVER_WET]++;
    if (correction == 1) merge_wetting_fronts(c);
    if (correction == 2) {
      double mass_before = mass_bal(c), mass_after;
      cross_layer_boundary(c);
      mass_after = mass_bal(c);
      if (fabs(mass_before - mass_after) > 100. * MBAL_ITERATIVE_TOLERANCE)
        *AET_demand_cm = *AET_demand_cm + (mass_before - mass_after);
    }
    if (correction == 3) {
      bottom_boundary_flux_cm += cross_domain_boundary(c);
      if (isnan(bottom_boundary_flux_cm)) bottom_boundary_flux_cm = 0.0;
    }
    if (correction == 4) {
      mass_change = 0.0;
      fix_dry_over_wet(c, &mass_change);
      *AET_demand_cm = *AET_demand_cm - mass_change;
    }
    correction = correction_type_surf(c);
  }

  /* refresh psi (only when psi > 1) and K of every front, :1824-1847 */
  for (wf = 1; wf != s->n + 1; wf++) {
    lgo_front* f = &s->f[wf - 1];
    const lgo_soil* sk = &p->soil[f->layer];
    double Se = Se_from_theta(f->theta, sk->theta_e, sk->theta_r);
    if (f->psi > 1.0) f->psi = lgo_h_from_Se(Se, sk->alpha, sk->m, sk->n);
    f->K = lgo_K_from_Se(Se, sk->Ksat, sk->m);
  }
  if (s->n == 1) { /* :1858-1863 */
    if (s->f[0].depth != cum[1]) s->f[0].depth = cum[1];
  }
  (void)mass_change;
  return bottom_boundary_flux_cm;
}

/* lgar_insert_water, lgar.cxx:2348-2484.  precip_timestep_cm is really a rate (quirk Q4). */
static double insert_water(ctx_t* c, double timestep_h, double AET_demand_cm, double free_drainage_subtimestep_cm,
                           double* ponded_depth_cm, double* volin_this_timestep, double precip_timestep_cm,
                           int wf_free_drainage_demand) {
  const lgo_state* s = c->s;
  const lgo_params* p = c->p;
  const double* cum = p->cum_thickness;
  int num_layers = p->num_layers;
  double f_p = 0.0, runoff = 0.0;
  double h_p = fmax(*ponded_depth_cm - precip_timestep_cm * timestep_h, 0.0);
  int fdi = wf_free_drainage_demand - 1;
  int number_of_wetting_fronts = s->n;
  int layer_num_fp = s->f[fdi].layer;
  const lgo_soil* sp = &p->soil[layer_num_fp];
  double Geff, Ksat_cm_per_h, current_mass, max_storage, ponded_depth_temp, free_drainage_demand, fp_cm;
  int k;

  if (number_of_wetting_fronts == num_layers) {
    Geff = 0.0;
    Ksat_cm_per_h = sp->Ksat;
  } else {
    double theta_below = s->f[fdi + 1].theta;
    Ksat_cm_per_h = sp->Ksat;
    Geff = lgo_Geff(p->use_closed_form_G, theta_below, sp->theta_e, sp, p->nint);
  }
  if (layer_num_fp == 1) {
    f_p = Ksat_cm_per_h * (1 + (Geff + h_p) / s->f[fdi].depth);
  } else {
    double bottom_sum = (s->f[fdi].depth - cum[layer_num_fp - 1]) / Ksat_cm_per_h;
    for (k = 1; k < layer_num_fp; k++) {
      double Ksat_cm_per_h_k = p->soil[layer_num_fp - k].Ksat;
      bottom_sum += (cum[layer_num_fp - k] - cum[layer_num_fp - (k + 1)]) / Ksat_cm_per_h_k;
    }
    f_p = (s->f[fdi].depth / bottom_sum) + ((Geff + h_p) * Ksat_cm_per_h / (s->f[fdi].depth));
  }
  current_mass = mass_bal(c);
  max_storage = 0.0;
  for (k = 1; k < num_layers + 1; k++) max_storage += p->soil[k].theta_e * (cum[k] - cum[k - 1]);
  if (f_p > (max_storage + free_drainage_subtimestep_cm + AET_demand_cm - current_mass) / timestep_h)
    f_p = (max_storage + free_drainage_subtimestep_cm + AET_demand_cm - current_mass) / timestep_h;

  free_drainage_demand = 0;
  ponded_depth_temp = *ponded_depth_cm - f_p * timestep_h - free_drainage_demand * 0;
  ponded_depth_temp = fmax(ponded_depth_temp, 0.0);
  fp_cm = f_p * timestep_h + free_drainage_demand / timestep_h;

  if (p->ponded_depth_max_cm > 0.0) {
    if (ponded_depth_temp < p->ponded_depth_max_cm) {
      runoff = 0.0;
      *volin_this_timestep = fmin(*ponded_depth_cm, fp_cm);
      *ponded_depth_cm = *ponded_depth_cm - *volin_this_timestep;
      return runoff;
    } else if (ponded_depth_temp > p->ponded_depth_max_cm) {
      runoff = ponded_depth_temp - p->ponded_depth_max_cm;
      *ponded_depth_cm = p->ponded_depth_max_cm;
      *volin_this_timestep = fp_cm;
      return runoff;
    }
    /* ponded_depth_temp == ponded_depth_max: the reference falls through and changes nothing */
  } else {
    *volin_this_timestep = fmin(*ponded_depth_cm, fp_cm);
    runoff = *ponded_depth_cm < fp_cm ? 0.0 : (*ponded_depth_cm - *volin_this_timestep);
    *ponded_depth_cm = 0.0;
  }
  return runoff;
}

/* lgar_calc_dry_depth, lgar.cxx:2577-2626 */
static double calc_dry_depth(ctx_t* c, double timestep_h, double* delta_theta) {
  const lgo_state* s = c->s;
  const lgo_params* p = c->p;
  int layer_num = s->f[0].layer;
  const lgo_soil* sp = &p->soil[layer_num];
  double theta1 = s->f[0].theta, theta_e = sp->theta_e, theta2 = theta_e;
  double tau, Geff, dry_depth;
  *delta_theta = theta_e - s->f[0].theta;
  tau = timestep_h * sp->Ksat / (theta_e - s->f[0].theta);
  Geff = lgo_Geff(p->use_closed_form_G, theta1, theta2, sp, p->nint);
  dry_depth = 0.5 * (tau + sqrt(tau * tau + 4.0 * tau * Geff));
  dry_depth = fmin(p->cum_thickness[layer_num], dry_depth);
  return dry_depth;
}

/* lgar_create_surficial_front, lgar.cxx:2490-2568 */
static void create_surficial_front(ctx_t* c, double* ponded_depth_cm, double* volin, double dry_depth, double theta1) {
  lgo_state* s = c->s;
  const lgo_params* p = c->p;
  const lgo_soil* sp = &p->soil[1];
  double theta_e = sp->theta_e, theta_r = sp->theta_r;
  double delta_theta = theta_e - theta1;
  double theta_new, Se;
  int lgo_path_note_ = (int)(c->s->n_path[LGO_PATH_SURFICIAL]++);
  (void)lgo_path_note_;
  if (dry_depth * delta_theta > (*ponded_depth_cm)) {
    *volin = *ponded_depth_cm;
    theta_new = fmin(theta1 + (*ponded_depth_cm) / dry_depth, theta_e);
    front_insert_first(c, dry_depth, theta_new, 1, 0);
    *ponded_depth_cm = 0.0;
  } else {
    *volin = dry_depth * delta_theta;
    *ponded_depth_cm -= dry_depth * delta_theta;
    theta_new = theta_e;
    front_insert_first(c, dry_depth, theta_e, 1, 0);
  }
  if (s->status & LGO_FAIL_OVERFLOW) return;
  Se = Se_from_theta(theta_new, theta_e, theta_r);
  s->f[0].psi = lgo_h_from_Se(Se, sp->alpha, sp->m, sp->n);
  s->f[0].K = lgo_K_from_Se(Se, sp->Ksat, sp->m) * 1.0; /* x frozen_factor (twice, quirk Q11) */
  s->f[0].dzdt = 0.0;
  if (s->n > 1) {
    int had_to_merge = 0;
    while ((s->n > 1) && (s->f[0].depth > s->f[1].depth) && (s->f[0].layer == s->f[1].layer) && !(s->f[1].to_bottom)) {
      merge_wetting_fronts(c);
      had_to_merge = 1;
      if (s->status & LGO_FAIL_LIST) break;
    }
    if (had_to_merge) cross_layer_boundary(c);
  }
}

/* lgar_dzdt_calc, lgar.cxx:2764-2945 */
static void dzdt_calc(ctx_t* c, double h_p, double subtimestep_h, int switch_caching, int cache_count, int new_front) {
  lgo_state* s = c->s;
  const lgo_params* p = c->p;
  const double* cum = p->cum_thickness;
  int num_layers = p->num_layers;
  int count_correct_for_crossing = 0;
  int i;
  for (i = 0; i < s->n; i++) {
    lgo_front* cur = &s->f[i];
    lgo_front* nxt;
    double dzdt = 0.0;
    int layer_num = cur->layer;
    double K_cm_per_h = cur->K;
    double depth_cm, Ksat_cm_per_h, theta1, theta2, bottom_sum, Geff, delta_theta, largest_K_s;
    const lgo_soil* sp = &p->soil[layer_num];
    int k, ii, is_next_to_bottom;
    if (K_cm_per_h < 0) {
      s->status |= LGO_FAIL_K_NEGATIVE;
      return;
    }
    depth_cm = cur->depth;
    Ksat_cm_per_h = sp->Ksat;
    if (i + 1 >= s->n) break; /* last front: nothing to do */
    nxt = &s->f[i + 1];
    theta1 = nxt->theta;
    theta2 = cur->theta;
    bottom_sum = 0.0;
    if (cur->to_bottom) {
      cur->dzdt = 0.0;
      continue;
    } else if (layer_num > 1) {
      bottom_sum += (cur->depth - cum[layer_num - 1]) / K_cm_per_h;
    }
    if (theta1 > theta2) {
      s->status |= LGO_FAIL_THETA_ORDER;
      return;
    }
    Geff = lgo_Geff(p->use_closed_form_G, theta1, theta2, sp, p->nint);
    delta_theta = cur->theta - nxt->theta;
    if (cur->layer == 1) {
      if (delta_theta > 0)
        dzdt = 1.0 / delta_theta * (Ksat_cm_per_h * (Geff + h_p) / cur->depth + cur->K);
      else
        dzdt = 0.0;
    } else {
      double denominator = bottom_sum, numerator;
      for (k = 1; k < layer_num; k++) { /* quirk Q3: K of layer (layer_num-k), thickness of layer k */
        const lgo_soil* sl = &p->soil[layer_num - k];
        double theta_prev_loc = theta_h(c, cur->psi, sl);
        double Se_prev_loc = Se_from_theta(theta_prev_loc, sl->theta_e, sl->theta_r);
        double K_cm_per_h_prev_loc = lgo_K_from_Se(Se_prev_loc, sl->Ksat, sl->m);
        denominator += (cum[k] - cum[k - 1]) / K_cm_per_h_prev_loc;
      }
      numerator = depth_cm;
      if (delta_theta > 0)
        dzdt = (1.0 / delta_theta) * ((numerator / denominator) + Ksat_cm_per_h * (Geff + h_p) / depth_cm);
      else
        dzdt = 0.0;
    }
    if ((dzdt == 0.0) && (cur->to_bottom == 0)) dzdt = 1e-9;
    if (dzdt > 1e4) dzdt = 1e4;
    largest_K_s = 0.0;
    for (ii = 1; ii <= num_layers; ii++) largest_K_s = fmax(p->soil[ii].Ksat, largest_K_s);
    if (dzdt > 100 * largest_K_s) dzdt = 100 * largest_K_s;
    if (dzdt < -100 * largest_K_s) dzdt = -100 * largest_K_s;
    is_next_to_bottom = nxt->to_bottom;
    if (is_next_to_bottom &&
        dzdt * subtimestep_h + cur->depth > FACTOR_LIMITS_LAYER_CROSSING_SPEED * (cum[cur->layer])) {
      count_correct_for_crossing++;
      dzdt = fabs(cur->depth - FACTOR_LIMITS_LAYER_CROSSING_SPEED * (cum[cur->layer])) / subtimestep_h +
             count_correct_for_crossing * DEPTH_AVOIDS_SAME_WF_DEPTH;
    }
    if (switch_caching) {
      if ((i + 1) != new_front) dzdt = dzdt * (cache_count);
    }
    cur->dzdt = dzdt;
  }
}

/* calc_aet, src/aet.cxx:17-77.  Depends on the head front only (quirk Q5). */
double lgo_calc_aet(double PET_timestep_cm, double time_step_h, double wilting_point_psi_cm,
                    double field_capacity_psi_cm, const lgo_front* head, const lgo_soil* sp) {
  double actual_ET_demand = 0.0;
  if (head->psi < 1.E6) {
    double theta_fc = lgo_theta_from_h(field_capacity_psi_cm, sp);
    double wp_head_theta = lgo_theta_from_h(wilting_point_psi_cm, sp);
    double theta_wp = (theta_fc - wp_head_theta) * 1 / 2 + wp_head_theta;
    double Se = Se_from_theta(theta_wp, sp->theta_e, sp->theta_r);
    double psi_wp_cm = lgo_h_from_Se(Se, sp->alpha, sp->m, sp->n);
    double h_ratio = 1.0 + pow(head->psi / psi_wp_cm, 3.0);
    actual_ET_demand = PET_timestep_cm * (1 / h_ratio) * time_step_h;
    if (actual_ET_demand < 0)
      actual_ET_demand = 0.0;
    else if (actual_ET_demand > (PET_timestep_cm * time_step_h))
      actual_ET_demand = PET_timestep_cm * time_step_h;
  }
  return actual_ET_demand;
}

/* calc_CR_Q, src/conceptual_reservoir.cxx:7-45 */
double lgo_calc_CR_Q(double subtimestep_h, double a_fast, double a_slow, double b_fast, double b_slow,
                     double frac_slow, double in_cm_per_h, double* fast, double* slow) {
  double input_slow = in_cm_per_h * frac_slow;
  double input_fast = in_cm_per_h - input_slow;
  double Q_fast, delta_fast, Q_slow, delta_slow;
  Q_fast = subtimestep_h * (a_fast * pow(*fast, b_fast));
  if (*fast < 0.01) Q_fast = 0.0;
  delta_fast = subtimestep_h * input_fast - Q_fast;
  if (*fast + delta_fast > 0.0) {
    *fast += delta_fast;
  } else {
    Q_fast = *fast + subtimestep_h * input_fast;
    *fast = 0.0;
  }
  Q_slow = subtimestep_h * (a_slow * pow(*slow, b_slow));
  if (*slow < 0.01) Q_slow = 0.0;
  delta_slow = subtimestep_h * input_slow - Q_slow;
  if (*slow + delta_slow > 0.0) {
    *slow += delta_slow;
  } else {
    Q_slow = *slow + subtimestep_h * input_slow;
    *slow = 0.0;
  }
  return Q_fast + Q_slow;
}

/* giuh_convolution_integral, giuh/giuh.c:12-39 */
double lgo_giuh(double runoff, int N, const double* ord, double* q) {
  double now;
  int i;
  q[N] = 0.0;
  for (i = 0; i < N; i++) q[i] += ord[i] * runoff;
  now = q[0];
  for (i = 1; i < N; i++) q[i - 1] = q[i];
  q[N - 1] = 0.0;
  return now;
}

/* ------------------------------------------------------------------------------------------ */
/* init: reference src/lgar.cxx:1024-1064 (InitializeWettingFronts) + :81-194, :984-1009       */
/* ------------------------------------------------------------------------------------------ */
void lgo_init_state(const lgo_params* p, lgo_state* s) {
  int layer;
  memset(s, 0, sizeof(*s));
  s->timestep_h = p->timestep_h;
  s->cache_count = 1;
  if (!p->invalid_soil_type) {
    for (layer = 1; layer <= p->num_layers; layer++) {
      const lgo_soil* sp = &p->soil[layer];
      lgo_front* f = &s->f[layer - 1];
      double Se;
      f->depth = p->cum_thickness[layer];
      f->theta = lgo_theta_from_h(p->initial_psi_cm, sp);
      f->layer = layer;
      f->to_bottom = 1;
      f->dzdt = 0.0;
      f->psi = p->initial_psi_cm;
      Se = Se_from_theta(f->theta, sp->theta_e, sp->theta_r);
      f->K = lgo_K_from_Se(Se, sp->Ksat, sp->m);
    }
    s->n = p->num_layers;
    s->volstart_cm = lgo_calc_mass_bal(p->cum_thickness, s->f, s->n);
  }
}

double lgo_global_balance(const lgo_state* m) { /* lgar.cxx:1185 (volchange_calib == 0) */
  return m->volstart_cm + m->volprecip_cm - m->volrunoff_cm - m->volAET_cm - m->volon_cm - m->volrech_cm -
         m->volend_cm + 0.0 - m->volrunoff_CR_cm - m->volCRend_cm;
}

/* ------------------------------------------------------------------------------------------ */
/* L3 timestep driver: BmiLGAR::Update(), reference src/bmi_lgar.cxx:133-895                   */
/* ------------------------------------------------------------------------------------------ */
int lgo_update(const lgo_params* p, lgo_state* s, double precipitation_mm_per_h, double PET_mm_per_h,
               double out[LGO_NSCALARS]) {
  ctx_t ctx = {p, s};
  ctx_t* c = &ctx;
  const double mm_to_cm = 0.1, mm_to_m = 0.001, cm_to_m = 0.01;
  int subcycles, num_layers = p->num_layers, cycle, i;
  double precip_timestep_cm = 0.0, PET_timestep_cm = 0.0, AET_timestep_cm = 0.0;
  double volend_timestep_cm, volCRend_timestep_cm, volin_timestep_cm = 0.0, volon_timestep_cm;
  double volrunoff_timestep_cm = 0.0, volrech_timestep_cm = 0.0, volrunoff_giuh_timestep_cm = 0.0;
  double volQ_timestep_cm = 0.0, volQ_CR_timestep_cm = 0.0;
  double precip_subtimestep_cm = 0.0, precip_subtimestep_cm_per_h, PET_subtimestep_cm = 0.0, PET_subtimestep_cm_per_h;
  double ponded_depth_subtimestep_cm, AET_subtimestep_cm = 0.0, volstart_subtimestep_cm, volend_subtimestep_cm;
  double volin_subtimestep_cm, volon_subtimestep_cm, volrunoff_subtimestep_cm, volrech_subtimestep_cm = 0.0;
  double precip_previous_subtimestep_cm = 0.0, volCRstart_subtimestep_cm, volCRend_subtimestep_cm;
  double subtimestep_h, wf_free_drain_dzdt = 0.0;
  int caching_at_start, switch_caching = 0;
  lgo_front old[LGO_MAX_FRONTS];
  int n_old;

  if (s->status & 0xFFFF) { /* a failed column is frozen (the reference process would be gone) */
    if (out) for (i = 0; i < LGO_NSCALARS; i++) out[i] = NAN;
    return s->status;
  }

  if (p->invalid_soil_type) { /* :145-179 */
    s->volprecip_cm += precipitation_mm_per_h * mm_to_cm;
    s->volin_cm = 0.0;
    s->volon_cm = 0.0;
    s->volend_cm = s->volstart_cm;
    s->volAET_cm = 0.0;
    s->volrech_cm = 0.0;
    s->volrunoff_cm += precipitation_mm_per_h * mm_to_cm;
    s->volQ_cm += precipitation_mm_per_h * mm_to_cm;
    s->volQ_CR_cm = 0.0;
    s->volPET_cm = 0.0;
    s->volrunoff_giuh_cm = 0.0;
    if (out) {
      for (i = 0; i < LGO_NSCALARS; i++) out[i] = 0.0;
      out[0] = precipitation_mm_per_h * mm_to_m;
      out[3] = precipitation_mm_per_h * mm_to_m;
      out[6] = precipitation_mm_per_h * mm_to_m;
    }
    s->timesteps++;
    return s->status;
  }

  volend_timestep_cm = mass_bal(c);
  volCRend_timestep_cm = s->CR_fast_storage_cm + s->CR_slow_storage_cm;
  volon_timestep_cm = s->volon_timestep_cm;
  volend_subtimestep_cm = volend_timestep_cm;
  volCRend_subtimestep_cm = volCRend_timestep_cm;
  subtimestep_h = s->timestep_h;

  /* adaptive sub-step, :260-270 (thresholds in mm/h, quirk Q7; timestep_h persistently overwritten) */
  if (p->adaptive_timestep) {
    subtimestep_h = p->forcing_resolution_h;
    if (precipitation_mm_per_h > 10.0 || volon_timestep_cm > 0.0)
      subtimestep_h = p->timestep_h;
    else if (precipitation_mm_per_h > 0.0)
      subtimestep_h = p->timestep_h * 2.0;
    subtimestep_h = fmin(subtimestep_h, p->forcing_resolution_h);
    s->timestep_h = subtimestep_h;
  }

  caching_at_start = s->cache_fluxes;
  if (!caching_at_start) wf_free_drain_dzdt = s->f[wf_free_drainage(c) - 1].dzdt;

  if (p->allow_flux_caching) { /* :282-307 */
    if (subtimestep_h == p->forcing_resolution_h) {
      if ((precipitation_mm_per_h < PRECIP_THRESHOLD_MM_PER_H) &&
          ((s->previous_AET / (s->previous_PET + PET_EPSILON)) < AET_PET_RATIO_THRESHOLD) &&
          (s->cache_count != NUM_TIMESTEPS_BEFORE_RESET_CACHE) && (volon_timestep_cm < VOLON_TIMESTEP_THRESHOLD_CM) &&
          (s->previous_recharge < BOTTOM_BDY_FLUX_THRESHOLD_CM))
        s->cache_fluxes = 1;
      if (wf_free_drain_dzdt > THRESHOLD_DZDT_CM_PER_H) s->cache_fluxes = 0;
      if (volon_timestep_cm >= VOLON_TIMESTEP_THRESHOLD_CM) s->cache_fluxes = 0;
      if ((s->cache_count == NUM_TIMESTEPS_BEFORE_RESET_CACHE) || (precipitation_mm_per_h >= PRECIP_THRESHOLD_MM_PER_H))
        s->cache_fluxes = 0;
    } else {
      s->cache_fluxes = 0;
    }
  }
  if (caching_at_start && !s->cache_fluxes) switch_caching = 1;
  if (s->cache_fluxes) s->cache_count++;
  if (s->cache_fluxes) s->n_path[LGO_PATH_CACHED_STEP]++;

  subcycles = (int)(p->forcing_resolution_h / s->timestep_h + 1.0e-08);

  if (precipitation_mm_per_h < 0.0) { /* :325-337 */
    precipitation_mm_per_h = 0.0;
    s->status |= LGO_WARN_NEG_FORCING;
  }
  if (PET_mm_per_h < 0.0) {
    PET_mm_per_h = 0.0;
    s->status |= LGO_WARN_NEG_FORCING;
  }
  if (p->PET_affects_precip) { /* :339-348 */
    if (precipitation_mm_per_h > PET_mm_per_h) {
      precipitation_mm_per_h = precipitation_mm_per_h - PET_mm_per_h;
      PET_mm_per_h = 0.0;
    } else {
      PET_mm_per_h = PET_mm_per_h - precipitation_mm_per_h;
      precipitation_mm_per_h = 0.0;
    }
  }

  for (cycle = 1; cycle <= subcycles; cycle++) { /* :356-805 */
    int top_near_sat = 0;
    double precip_for_CR_subtimestep_cm_per_h = 0.0, ponded_flux_for_CR = 0.0;
    double temp_rch = 0.0, free_drainage_subtimestep_cm = 0.0, free_drainage_for_CR = 0.0;
    double volQ_CR_subtimestep_cm, local_mb;

    s->timesteps++;
    precip_subtimestep_cm_per_h = precipitation_mm_per_h * mm_to_cm;
    n_old = s->n;
    memcpy(old, s->f, sizeof(lgo_front) * (size_t)s->n); /* state_previous = listCopy(head) :370-374 */

    if (s->runoff_in_prev_step) { /* preferential flow, :379-386 (quirk Q9) */
      double total = precip_subtimestep_cm_per_h;
      precip_for_CR_subtimestep_cm_per_h = p->frac_to_CR * total;
      precip_subtimestep_cm_per_h = (1.0 - p->frac_to_CR) * total;
      ponded_flux_for_CR = volon_timestep_cm / subtimestep_h * p->frac_to_CR;
      volon_timestep_cm = volon_timestep_cm * (1 - p->frac_to_CR);
    }

    AET_subtimestep_cm = 0.0;
    volin_subtimestep_cm = 0.0;
    volon_subtimestep_cm = 0.0;
    volrunoff_subtimestep_cm = 0.0;
    volrech_subtimestep_cm = 0.0;
    PET_subtimestep_cm_per_h = PET_mm_per_h * mm_to_cm;
    ponded_depth_subtimestep_cm = precip_subtimestep_cm_per_h * subtimestep_h;
    ponded_depth_subtimestep_cm += volon_timestep_cm;
    precip_subtimestep_cm = precip_subtimestep_cm_per_h * subtimestep_h;
    PET_subtimestep_cm = PET_subtimestep_cm_per_h * subtimestep_h;
    volstart_subtimestep_cm = mass_bal(c);

    if (!s->cache_fluxes) { /* :418-681 */
      double min_storage = 0.0;
      double mass_used_to_check_impossible_storages = mass_bal(c);
      int k, wf_free_drainage_demand, iter_chk, create_front, is_top_wf_saturated, new_front;
      double min_water_possible_for_FD_WF, storage_in_FD_WF, mass_correction_for_cached_fd = 0.0;
      double theta_e, theta_above_which_precip_contribs_to_GW, delta_theta, dry_depth;

      s->n_active_substeps++;
      s->n_path[LGO_PATH_ACTIVE_STEP]++;
      for (k = 1; k < num_layers + 1; k++)
        min_storage += p->soil[k].theta_r * (p->cum_thickness[k] - p->cum_thickness[k - 1]);
      wf_free_drainage_demand = wf_free_drainage(c);
      min_water_possible_for_FD_WF = min_water_fd(c, wf_free_drainage_demand);
      storage_in_FD_WF = storage_fd(c, wf_free_drainage_demand);

      if (switch_caching) { /* :436-441 (adds a depth to a rate, quirk Q6) */
        PET_subtimestep_cm_per_h += s->accumulated_PET;
        s->accumulated_PET = 0.0;
        mass_correction_for_cached_fd += s->accumulated_free_drainage;
        s->accumulated_free_drainage = 0.0;
      }

      if (p->free_drainage_enabled) { /* :443-461 */
        const lgo_front* front = &s->f[s->n - 1];
        free_drainage_subtimestep_cm += subtimestep_h * front->K;
        iter_chk = 0;
        while ((mass_used_to_check_impossible_storages - free_drainage_subtimestep_cm < min_storage) ||
               (storage_in_FD_WF - free_drainage_subtimestep_cm < min_water_possible_for_FD_WF)) {
          free_drainage_subtimestep_cm *= 0.5;
          if (iter_chk > 5) {
            free_drainage_subtimestep_cm = 0.0;
            break;
          }
          iter_chk++;
        }
        if (free_drainage_subtimestep_cm < 1.E-7) free_drainage_subtimestep_cm = 0.0;
        if (front->psi > 1.E6) free_drainage_subtimestep_cm = 0.0;
      }

      precip_previous_subtimestep_cm = s->precip_previous_timestep_cm;

      if (PET_subtimestep_cm_per_h > 0.0) /* :481-485 */
        AET_subtimestep_cm = lgo_calc_aet(PET_subtimestep_cm_per_h, subtimestep_h, p->wilting_point_psi_cm,
                                          p->field_capacity_psi_cm, &s->f[0], &p->soil[s->f[0].layer]);
      s->n_theta_calls += (PET_subtimestep_cm_per_h > 0.0 && s->f[0].psi < 1.E6) ? 2 : 0;

      iter_chk = 0; /* :487-495 */
      while ((mass_used_to_check_impossible_storages - AET_subtimestep_cm < min_storage) ||
             (storage_in_FD_WF - AET_subtimestep_cm < min_water_possible_for_FD_WF)) {
        AET_subtimestep_cm *= 0.5;
        if (iter_chk > 5) {
          AET_subtimestep_cm = 0.0;
          break;
        }
        iter_chk++;
      }
      iter_chk = 0; /* :497-510 */
      while (((mass_used_to_check_impossible_storages - AET_subtimestep_cm - free_drainage_subtimestep_cm -
               mass_correction_for_cached_fd) < min_storage) ||
             ((storage_in_FD_WF - AET_subtimestep_cm - free_drainage_subtimestep_cm - mass_correction_for_cached_fd) <
              min_water_possible_for_FD_WF)) {
        AET_subtimestep_cm *= 0.5;
        free_drainage_subtimestep_cm *= 0.5;
        mass_correction_for_cached_fd *= 0.5;
        if (iter_chk > 5) {
          AET_subtimestep_cm = 0.0;
          free_drainage_subtimestep_cm = 0.0;
          mass_correction_for_cached_fd = 0.0;
          break;
        }
        iter_chk++;
      }

      precip_timestep_cm += precip_subtimestep_cm + precip_for_CR_subtimestep_cm_per_h * subtimestep_h;
      PET_timestep_cm += fmax(PET_subtimestep_cm, 0.0);
      volon_timestep_cm = fmax(volon_timestep_cm, 0.0);
      volon_timestep_cm = volon_timestep_cm > SMALL_EPS ? volon_timestep_cm : 0.0;

      /* should a new surficial front be created? :522-535 (quirk Q8) */
      theta_e = p->soil[s->f[0].layer].theta_e;
      is_top_wf_saturated = (s->f[0].theta + SMALL_EPS) >= theta_e ? 1 : 0;
      theta_above_which_precip_contribs_to_GW = theta_e * p->spf_factor;
      top_near_sat = s->f[0].theta > theta_above_which_precip_contribs_to_GW ? 1 : 0;
      create_front =
          (precip_previous_subtimestep_cm == 0.0 && precip_subtimestep_cm > 0.0 && volon_timestep_cm == 0) ||
          ((precip_subtimestep_cm > 0.0 || volon_timestep_cm > 0) && (s->n == num_layers));
      if (is_top_wf_saturated || volon_timestep_cm > 0.0) create_front = 0;

      if (create_front) { /* :547-595 */
        double temp_pd = 0.0;
        temp_rch = move_wetting_fronts(c, subtimestep_h, &free_drainage_subtimestep_cm, &temp_pd, wf_free_drainage_demand,
                                       volend_subtimestep_cm, mass_correction_for_cached_fd, &AET_subtimestep_cm, old, n_old);
        dry_depth = calc_dry_depth(c, subtimestep_h, &delta_theta);
        create_surficial_front(c, &ponded_depth_subtimestep_cm, &volin_subtimestep_cm, dry_depth, s->f[0].theta);
        n_old = s->n;
        memcpy(old, s->f, sizeof(lgo_front) * (size_t)s->n);
      }

      if (ponded_depth_subtimestep_cm > 0 && !create_front) { /* :601-618 */
        volrunoff_subtimestep_cm =
            insert_water(c, subtimestep_h, AET_subtimestep_cm, free_drainage_subtimestep_cm, &ponded_depth_subtimestep_cm,
                         &volin_subtimestep_cm, precip_subtimestep_cm_per_h, wf_free_drainage_demand);
        volrech_subtimestep_cm = volin_subtimestep_cm;
        volon_subtimestep_cm = ponded_depth_subtimestep_cm;
        if (volrunoff_subtimestep_cm < 0) {
          s->status |= LGO_FAIL_NEG_RUNOFF;
          break;
        }
      } else { /* :619-632 */
        if (ponded_depth_subtimestep_cm < p->ponded_depth_max_cm) {
          volon_subtimestep_cm = ponded_depth_subtimestep_cm;
          ponded_depth_subtimestep_cm = 0.0;
          volrunoff_subtimestep_cm = 0.0;
        } else {
          volrunoff_subtimestep_cm = (ponded_depth_subtimestep_cm - p->ponded_depth_max_cm);
          volon_subtimestep_cm = p->ponded_depth_max_cm;
          ponded_depth_subtimestep_cm = p->ponded_depth_max_cm;
        }
      }

      if (!create_front) { /* :638-652 */
        double volin_subtimestep_cm_temp = volin_subtimestep_cm;
        temp_rch = move_wetting_fronts(c, subtimestep_h, &free_drainage_subtimestep_cm, &volin_subtimestep_cm,
                                       wf_free_drainage_demand, volend_subtimestep_cm, mass_correction_for_cached_fd,
                                       &AET_subtimestep_cm, old, n_old);
        volrech_subtimestep_cm = volin_subtimestep_cm;
        volin_subtimestep_cm = volin_subtimestep_cm_temp;
      }

      clean_redundant_fronts(c); /* :654 */

      new_front = -1; /* :658-666 */
      if (switch_caching) {
        if (create_front) new_front = 1;
      }
      dzdt_calc(c, ponded_depth_subtimestep_cm, subtimestep_h, switch_caching, s->cache_count, new_front);
      if (s->status & 0xFFFF) break;
      if (switch_caching) s->cache_count = 1;

      volrech_subtimestep_cm = temp_rch; /* :672-679 */
      if (!p->free_drainage_to_CR) {
        volrech_subtimestep_cm += free_drainage_subtimestep_cm;
      } else {
        free_drainage_for_CR += free_drainage_subtimestep_cm + volrech_subtimestep_cm;
        volrech_subtimestep_cm = 0.0;
      }
    } else { /* cached fluxes, :683-695 (quirk Q6) */
      PET_timestep_cm += fmax(PET_subtimestep_cm, 0.0);
      s->accumulated_PET += PET_timestep_cm;
      if (!p->free_drainage_to_CR)
        volrech_subtimestep_cm += s->previous_recharge;
      else
        free_drainage_for_CR += s->previous_recharge;
      s->accumulated_free_drainage += s->previous_recharge;
    }

    volend_subtimestep_cm = mass_bal(c); /* :697-707 */
    volend_timestep_cm = volend_subtimestep_cm;
    s->precip_previous_timestep_cm = precip_subtimestep_cm;
    volCRstart_subtimestep_cm = s->CR_fast_storage_cm + s->CR_slow_storage_cm;
    volQ_CR_subtimestep_cm = lgo_calc_CR_Q(subtimestep_h, p->a, p->a_slow, p->b, p->b_slow, p->frac_slow,
                                           precip_for_CR_subtimestep_cm_per_h + ponded_flux_for_CR +
                                               free_drainage_for_CR / subtimestep_h,
                                           &s->CR_fast_storage_cm, &s->CR_slow_storage_cm);
    s->volrunoff_CR_cm += volQ_CR_subtimestep_cm;
    volQ_CR_timestep_cm += volQ_CR_subtimestep_cm;
    volCRend_subtimestep_cm = s->CR_fast_storage_cm + s->CR_slow_storage_cm;
    volCRend_timestep_cm = volCRend_subtimestep_cm;

    if ((volrunoff_subtimestep_cm > SMALL_EPS) || (top_near_sat)) /* :710-715 */
      s->runoff_in_prev_step = 1;
    else
      s->runoff_in_prev_step = 0;

    precip_subtimestep_cm += precip_for_CR_subtimestep_cm_per_h * subtimestep_h; /* :718 */

    local_mb = volstart_subtimestep_cm + precip_subtimestep_cm + volon_timestep_cm - volrunoff_subtimestep_cm -
               volQ_CR_subtimestep_cm - volCRend_subtimestep_cm + volCRstart_subtimestep_cm - AET_subtimestep_cm -
               volon_subtimestep_cm - volrech_subtimestep_cm - volend_subtimestep_cm; /* :723-724 */

    volin_timestep_cm += volin_subtimestep_cm; /* :731-737 */
    volrech_timestep_cm += volrech_subtimestep_cm;
    volrunoff_timestep_cm += volrunoff_subtimestep_cm;
    AET_timestep_cm += AET_subtimestep_cm;
    volon_timestep_cm = volon_subtimestep_cm;

    s->local_mass_balance = local_mb;
    if (fabs(local_mb) > p->mbal_tol || isinf(local_mb)) { /* :745-788 */
      s->status |= LGO_FAIL_MBAL;
      break;
    }
    if (!(s->f[0].depth > 0.0)) { /* :795 */
      s->status |= LGO_FAIL_TOP_DEPTH;
      break;
    }
    if (s->status & 0xFFFF) break;
  }

  /* GIUH once per forcing step, input in cm (quirk Q12), :808-812 */
  volrunoff_giuh_timestep_cm = lgo_giuh(volrunoff_timestep_cm + volQ_CR_timestep_cm, p->num_giuh, p->giuh, s->giuh_queue);
  volQ_timestep_cm = volrunoff_giuh_timestep_cm;

  { /* :824-834 */
    int to_bottom_count = 0;
    for (i = 0; i < s->n; i++) {
      if (s->f[i].to_bottom) to_bottom_count++;
      if (to_bottom_count > num_layers) s->status |= LGO_FAIL_TO_BOTTOM;
    }
  }

  s->volon_timestep_cm = volon_timestep_cm; /* :848 */
  s->previous_AET = AET_subtimestep_cm;     /* :860-864 */
  s->previous_PET = PET_subtimestep_cm;
  if (!s->cache_fluxes) s->previous_recharge = volrech_subtimestep_cm;

  s->volprecip_cm += precip_timestep_cm; /* :867-878 */
  s->volin_cm += volin_timestep_cm;
  s->volon_cm = volon_timestep_cm;
  s->volend_cm = volend_timestep_cm;
  s->volCRend_cm = volCRend_timestep_cm;
  s->volAET_cm += AET_timestep_cm;
  s->volrech_cm += volrech_timestep_cm;
  s->volrunoff_cm += volrunoff_timestep_cm;
  s->volQ_cm += volQ_timestep_cm;
  s->volQ_CR_cm += volQ_CR_timestep_cm;
  s->volPET_cm += PET_timestep_cm;
  s->volrunoff_giuh_cm += volrunoff_giuh_timestep_cm;

  if (out) { /* :882-893 */
    out[0] = precip_timestep_cm * cm_to_m;
    out[1] = PET_timestep_cm * cm_to_m;
    out[2] = AET_timestep_cm * cm_to_m;
    out[3] = volrunoff_timestep_cm * cm_to_m;
    out[4] = volrunoff_giuh_timestep_cm * cm_to_m;
    out[5] = volend_timestep_cm * cm_to_m;
    out[6] = volQ_timestep_cm * cm_to_m;
    out[7] = volin_timestep_cm * cm_to_m;
    out[8] = volrech_timestep_cm * cm_to_m;
    out[9] = volQ_CR_timestep_cm * cm_to_m;
    out[10] = s->local_mass_balance * cm_to_m;
    out[11] = volCRend_timestep_cm * cm_to_m;
  }
  (void)volCRend_subtimestep_cm;
  return s->status;
}

/* same record layout as oracle/ref_harness.cpp:lgref_run_series */
int lgo_run_series(const lgo_params* p, lgo_state* s, int nsteps, const double* precip, const double* pet,
                   double* scalars, int* nfronts, int cap, double* fronts, int* flags) {
  int t, i;
  for (t = 0; t < nsteps; t++) {
    int st = lgo_update(p, s, precip[t], pet[t], scalars ? scalars + (size_t)LGO_NSCALARS * t : (double*)0);
    if (st & 0xFFFF) return t;
    if (nfronts) nfronts[t] = s->n;
    if (fronts) {
      for (i = 0; i < s->n && i < cap; i++) {
        double* q = fronts + ((size_t)t * cap + i) * 5;
        q[0] = s->f[i].depth;
        q[1] = s->f[i].theta;
        q[2] = s->f[i].psi;
        q[3] = s->f[i].K;
        q[4] = s->f[i].dzdt;
        flags[(size_t)t * cap + i] = s->f[i].layer | (s->f[i].to_bottom << 8);
      }
    }
  }
  return nsteps;
}
