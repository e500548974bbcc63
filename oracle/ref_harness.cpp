// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// Thin C-ABI driver around the UNMODIFIED reference sources in /root/reference (compiled where
// they lie by oracle/Makefile into oracle/_ref/liblgar_ref.so).  It owns BmiLGAR objects, feeds
// them forcing through the reference's own BMI verbs (SetValue/Update/GetValue,
// reference src/bmi_main_lgar.cxx:116-152) and exposes the raw wetting-front list from
// BmiLGAR::get_model()->head (reference src/bmi_lgar.cxx:905) in full double precision, because
// the reference's own CSV writer only prints 6 decimals (src/bmi_main_lgar.cxx:273).
//
// Nothing of the reference is copied here: this file only *calls* its public interface.
#include <csetjmp>
#include <csignal>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <chrono>

#include "bmi.hxx"
#include "all.hxx"
#include "bmi_lgar.hxx"

static sigjmp_buf g_jmp;
static volatile int g_armed = 0;

// The reference calls exit() on a few numerical failures (src/lgar.cxx:2695,2710,2852).  The
// library is linked -Wl,-Bsymbolic-functions with its own `exit` (oracle/ref_exit_hook.c), which
// lands here, so one bad column cannot take the test process down.
extern "C" void lgref_exit_hook(int code) {
  if (g_armed) siglongjmp(g_jmp, 2);
  _Exit(code ? code : 3);
}

static void on_abort(int) {
  if (g_armed) siglongjmp(g_jmp, 1);
  _Exit(4);
}

struct RefCol {
  BmiLGAR model;
  bool dead = false;
};

static const char* kScalarNames[12] = {
    "precipitation", "potential_evapotranspiration", "actual_evapotranspiration",
    "surface_runoff", "giuh_runoff", "soil_storage", "total_discharge", "infiltration",
    "percolation", "conceptual_reservoir_to_stream_discharge", "mass_balance",
    "conceptual_reservoir_storage"};

extern "C" {

void* lgref_create(const char* config_path) {
  struct sigaction sa;
  memset(&sa, 0, sizeof(sa));
  sa.sa_handler = on_abort;
  sa.sa_flags = SA_NODEFER;
  sigaction(SIGABRT, &sa, nullptr);
  RefCol* c = new RefCol();
  g_armed = 1;
  if (sigsetjmp(g_jmp, 1) != 0) {
    g_armed = 0;
    return nullptr;  // leaked on purpose: the object is in an undefined state
  }
  try {
    c->model.Initialize(std::string(config_path));
  } catch (std::exception& e) {
    fprintf(stderr, "lgref_create: %s\n", e.what());
    g_armed = 0;
    return nullptr;
  }
  g_armed = 0;
  return c;
}

// one forcing step. returns 0 ok, 1 = abort()/assert, 2 = exit()
int lgref_update(void* h, double precip_mm_h, double pet_mm_h) {
  RefCol* c = (RefCol*)h;
  if (c->dead) return 1;
  g_armed = 1;
  int rc = sigsetjmp(g_jmp, 1);
  if (rc != 0) {
    g_armed = 0;
    c->dead = true;
    return rc;
  }
  c->model.SetValue("precipitation_rate", &precip_mm_h);
  c->model.SetValue("potential_evapotranspiration_rate", &pet_mm_h);
  c->model.Update();
  g_armed = 0;
  return 0;
}

// SetValue / GetValuePtr of a scalar double variable by name (calibratable parameters, src/bmi_lgar.cxx:1372-1421,1455).
// returns 0 ok, 1 = unknown variable
int lgref_set_scalar(void* h, const char* name, double value) {
  RefCol* c = (RefCol*)h;
  try {
    c->model.SetValue(std::string(name), &value);
  } catch (std::exception&) {
    return 1;
  }
  return 0;
}
// SetValue of an array variable (soil_temperature_profile of the SFT coupling): copies GetVarNbytes(name) bytes
int lgref_set_array(void* h, const char* name, double* values) {
  RefCol* c = (RefCol*)h;
  try {
    c->model.SetValue(std::string(name), values);
  } catch (std::exception&) {
    return 1;
  }
  return 0;
}
int lgref_get_scalar(void* h, const char* name, double* value) {
  RefCol* c = (RefCol*)h;
  try {
    *value = *static_cast<double*>(c->model.GetValuePtr(std::string(name)));
  } catch (std::exception&) {
    return 1;
  }
  return 0;
}

// the 11 BMI output scalars (order of src/bmi_main_lgar.cxx:63-73) + conceptual_reservoir_storage, in m
void lgref_get_scalars(void* h, double* out12) {
  RefCol* c = (RefCol*)h;
  for (int i = 0; i < 11; i++) c->model.GetValue(kScalarNames[i], &out12[i]);
  // "conceptual_reservoir_storage" has no grid (GetVarGrid -> -1, src/bmi_lgar.cxx:1112-1154), so
  // GetValue copies 0 bytes; it is only reachable through GetValuePtr (src/bmi_lgar.cxx:1338).
  out12[11] = *static_cast<double*>(c->model.GetValuePtr(kScalarNames[11]));
}

int lgref_num_fronts(void* h) {
  RefCol* c = (RefCol*)h;
  int n = 0;
  for (wetting_front* f = c->model.get_model()->head; f; f = f->next) n++;
  return n;
}

// raw front list; returns n (copies at most cap)
int lgref_get_fronts(void* h, int cap, double* depth_cm, double* theta, double* psi_cm,
                     double* K_cm_h, double* dzdt_cm_h, int* layer, int* to_bottom) {
  RefCol* c = (RefCol*)h;
  int n = 0;
  for (wetting_front* f = c->model.get_model()->head; f; f = f->next) {
    if (n < cap) {
      depth_cm[n] = f->depth_cm; theta[n] = f->theta; psi_cm[n] = f->psi_cm;
      K_cm_h[n] = f->K_cm_per_h; dzdt_cm_h[n] = f->dzdt_cm_per_h;
      layer[n] = f->layer_num; to_bottom[n] = f->to_bottom ? 1 : 0;
    }
    n++;
  }
  return n;
}

// BMI per-front outputs through GetValue (theta, depth in m); returns n
int lgref_get_bmi_fronts(void* h, int cap, double* theta, double* depth_m) {
  RefCol* c = (RefCol*)h;
  int n = 0;
  c->model.GetValue("soil_num_wetting_fronts", &n);
  if (n > cap) return n;
  c->model.GetValue("soil_moisture_wetting_fronts", theta);
  c->model.GetValue("soil_depth_wetting_fronts", depth_m);
  return n;
}

// persistent scalars of the model state (for debugging parity of the carried state)
void lgref_get_state(void* h, double* out16) {
  RefCol* c = (RefCol*)h;
  model_state* s = c->model.get_model();
  out16[0] = s->lgar_mass_balance.volon_timestep_cm;
  out16[1] = s->lgar_bmi_params.precip_previous_timestep_cm;
  out16[2] = s->lgar_mass_balance.CR_fast_storage_cm;
  out16[3] = s->lgar_mass_balance.CR_slow_storage_cm;
  out16[4] = s->lgar_mass_balance.previous_AET;
  out16[5] = s->lgar_mass_balance.previous_PET;
  out16[6] = s->lgar_mass_balance.previous_recharge;
  out16[7] = s->lgar_mass_balance.accumulated_PET;
  out16[8] = s->lgar_mass_balance.accumulated_free_drainage;
  out16[9] = s->lgar_bmi_params.timestep_h;
  out16[10] = s->lgar_mass_balance.cache_fluxes ? 1.0 : 0.0;
  out16[11] = (double)s->lgar_bmi_params.cache_count;
  out16[12] = s->lgar_bmi_params.runoff_in_prev_step ? 1.0 : 0.0;
  out16[13] = s->lgar_bmi_params.time_s;
  out16[14] = (double)s->lgar_bmi_params.timesteps;
  out16[15] = s->lgar_mass_balance.local_mass_balance;
}

// cumulative volumes for the global mass balance (src/lgar.cxx:1162-1185), cm
void lgref_get_cumulative(void* h, double* out16) {
  RefCol* c = (RefCol*)h;
  lgar_mass_balance_variables& m = c->model.get_model()->lgar_mass_balance;
  out16[0] = m.volstart_cm; out16[1] = m.volprecip_cm; out16[2] = m.volin_cm;
  out16[3] = m.volend_cm; out16[4] = m.volon_cm; out16[5] = m.volrunoff_cm;
  out16[6] = m.volrunoff_giuh_cm; out16[7] = m.volrech_cm; out16[8] = m.volAET_cm;
  out16[9] = m.volPET_cm; out16[10] = m.volrunoff_CR_cm; out16[11] = m.volCRend_cm;
  out16[12] = m.volQ_cm; out16[13] = m.volchange_calib_cm; out16[14] = m.volQ_CR_cm;
  out16[15] = m.volstart_cm + m.volprecip_cm - m.volrunoff_cm - m.volAET_cm - m.volon_cm -
              m.volrech_cm - m.volend_cm + m.volchange_calib_cm - m.volrunoff_CR_cm - m.volCRend_cm;
}

// the parsed soil table row (reference parser quirks included, src/lgar.cxx:2674-2758)
// out10 = theta_r, theta_e, alpha, n, m, lambda, psib, h_min, Ksat, theta_wp
void lgref_get_soil(void* h, int soil, double* out10) {
  RefCol* c = (RefCol*)h;
  soil_properties_& p = c->model.get_model()->soil_properties[soil];
  out10[0] = p.theta_r; out10[1] = p.theta_e; out10[2] = p.vg_alpha_per_cm; out10[3] = p.vg_n;
  out10[4] = p.vg_m; out10[5] = p.bc_lambda; out10[6] = p.bc_psib_cm; out10[7] = p.h_min_cm;
  out10[8] = p.Ksat_cm_per_h; out10[9] = p.theta_wp;
}

int lgref_num_giuh(void* h) { return ((RefCol*)h)->model.get_model()->lgar_bmi_params.num_giuh_ordinates; }

// Run nsteps forcing steps for one column.  scalars (nullable): [nsteps][12]; nfronts (nullable):
// [nsteps]; fronts (nullable): [nsteps][cap][5] = depth,theta,psi,K,dzdt ; flags [nsteps][cap] =
// layer | to_bottom<<8.  Returns number of steps completed (< nsteps when the reference aborted).
int lgref_run_series(void* h, int nsteps, const double* precip, const double* pet, double* scalars,
                     int* nfronts, int cap, double* fronts, int* flags) {
  RefCol* c = (RefCol*)h;
  for (int t = 0; t < nsteps; t++) {
    int rc = lgref_update(h, precip[t], pet[t]);
    if (rc != 0) return t;
    if (scalars) lgref_get_scalars(h, scalars + (size_t)12 * t);
    int n = 0;
    if (nfronts || fronts) {
      for (wetting_front* f = c->model.get_model()->head; f; f = f->next) {
        if (fronts && n < cap) {
          double* p = fronts + ((size_t)t * cap + n) * 5;
          p[0] = f->depth_cm; p[1] = f->theta; p[2] = f->psi_cm; p[3] = f->K_cm_per_h; p[4] = f->dzdt_cm_per_h;
          flags[(size_t)t * cap + n] = f->layer_num | ((f->to_bottom ? 1 : 0) << 8);
        }
        n++;
      }
      if (nfronts) nfronts[t] = n;
    }
  }
  return nsteps;
}

// CPU-baseline timing helper: many columns, one shared forcing series with a per-column circular
// shift and precipitation multiplier; SetValue x2 -> Update -> GetValue(total_discharge) per step,
// no file I/O.  Returns wall seconds spent in the Update loop; *sum_q accumulates discharge so the
// loop cannot be optimised away.
double lgref_time_columns(void** hs, int ncol, int nsteps, const double* precip, const double* pet,
                          int series_len, const int* shift, const double* pmult, double* sum_q) {
  auto t0 = std::chrono::steady_clock::now();
  double acc = 0.0;
  for (int t = 0; t < nsteps; t++) {
    for (int i = 0; i < ncol; i++) {
      RefCol* c = (RefCol*)hs[i];
      if (c->dead) continue;
      int k = (t + shift[i]) % series_len;
      double p = precip[k] * pmult[i], e = pet[k];
      if (lgref_update(c, p, e) != 0) continue;
      double q;
      c->model.GetValue("total_discharge", &q);
      acc += q;
    }
  }
  auto t1 = std::chrono::steady_clock::now();
  *sum_q = acc;
  return std::chrono::duration<double>(t1 - t0).count();
}

// Same loop with a per-column forcing series laid out [nsteps][ncol] (mm/h).
double lgref_time_columns_series(void** hs, int ncol, int nsteps, const double* precip, const double* pet, double* sum_q) {
  auto t0 = std::chrono::steady_clock::now();
  double acc = 0.0;
  for (int t = 0; t < nsteps; t++) {
    for (int i = 0; i < ncol; i++) {
      RefCol* c = (RefCol*)hs[i];
      if (c->dead) continue;
      if (lgref_update(c, precip[(size_t)t * ncol + i], pet[(size_t)t * ncol + i]) != 0) continue;
      double q;
      c->model.GetValue("total_discharge", &q);
      acc += q;
    }
  }
  auto t1 = std::chrono::steady_clock::now();
  *sum_q = acc;
  return std::chrono::duration<double>(t1 - t0).count();
}

void lgref_destroy(void* h) {
  RefCol* c = (RefCol*)h;
  if (!c) return;
  if (!c->dead) {
    // Finalize() prints the global mass balance to stdout (src/bmi_lgar.cxx:1081); silence it.
    FILE* saved = stdout;
    FILE* nul = fopen("/dev/null", "w");
    if (nul) stdout = nul;
    c->model.Finalize();
    if (nul) { fclose(nul); }
    stdout = saved;
    delete c;
  }
}

}  // extern "C"
