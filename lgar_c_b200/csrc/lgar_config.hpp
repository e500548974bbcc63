// lgar_c_b200/csrc/lgar_config.hpp -- host-only (no CUDA) readers for the reference's input files
// and the derivation of the per-class constants the kernels consume.
//
//   parse_config        mirrors InitFromConfigFile        (reference src/lgar.cxx:231-1016)
//   read_soil_file      mirrors lgar_read_vG_param_file   (reference src/lgar.cxx:2674-2758, quirk Q2)
//   make_engine_config  GIUH re-sampling etc.             (reference src/lgar.cxx:918-929,992)
//   make_class          soil_properties_ rows -> LgarColumnClass, initial fronts
//                                                          (reference src/lgar.cxx:1024-1064)
#ifndef LGAR_CONFIG_HPP
#define LGAR_CONFIG_HPP

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "lgar_types.h"

namespace lgarhost {

#define MAX_NUM_SOIL_TYPES 25  // reference include/all.hxx:39


struct SoilRow {  // reference struct soil_properties_, include/all.hxx:66-79
  double theta_r = 0, theta_e = 0, alpha = 0, n = 0, m = 0, lambda = 0, psib = 0, h_min = 0, Ksat = 0, theta_wp = 0;
};

struct ParsedConfig {
  std::vector<double> layer_thickness;
  std::vector<int> layer_soil_type;
  std::vector<double> giuh_hourly;
  std::vector<double> soil_z;  // depths of the soil temperature cells (SFT coupling), in the unit of the layer thicknesses
  std::string soil_params_file;
  double initial_psi = 0, wilting_point_psi = 0, field_capacity_psi = 0;
  double a = 0, b = 0, frac_to_CR = 0, a_slow = 0, b_slow = 0, frac_slow = 0, spf_factor = 0.98;
  double mbal_tol = 1.E1, timestep_h = 0, endtime_s = 0, forcing_resolution_h = 0, ponded_depth_max = 0;
  int num_soil_types = MAX_NUM_SOIL_TYPES;
  bool use_closed_form_G = false, adaptive_timestep = false, PET_affects_precip = false, allow_flux_caching = false;
  bool log_mode = false, free_drainage_enabled = false, free_drainage_to_CR = false, sft_coupled = false, calib_params = false;
};

inline double theta_from_h(double h, const SoilRow& s) {  // soil_funcs.cxx:147-150
  return (1.0 / (pow(1.0 + pow(s.alpha * h, s.n), s.m)) * (s.theta_e - s.theta_r) + s.theta_r);
}
inline double K_from_Se(double Se, double Ks, double m) {  // soil_funcs.cxx:164-167
  return (Ks * sqrt(Se) * pow(1.0 - pow(1.0 - pow(Se, 1.0 / m), m), 2.0));
}
inline double h_from_Se(double Se, double alpha, double m, double n) {  // soil_funcs.cxx:172-179
  double r = 1.0 / alpha * pow(pow(Se, -1.0 / m) - 1.0, 1.0 / n);
  return r > 1.E20 ? 1.E20 : r;
}

// ReadVectorData, lgar.cxx:1073-1100
inline std::vector<double> read_vector(const std::string& key) {
  std::vector<double> value;
  std::string z1 = key;
  const std::string delim = ",";
  if (z1.find(delim) == std::string::npos) {
    value.push_back(std::stod(z1));
  } else {
    while (z1.find(delim) != std::string::npos) {
      size_t pos = z1.find(delim);
      value.push_back(std::stod(z1.substr(0, pos)));
      z1.erase(0, pos + delim.length());
      if (z1.find(delim) == std::string::npos) value.push_back(std::stod(z1));
    }
  }
  return value;
}

inline bool parse_bool(const std::string& key, const std::string& v, bool allow01) {
  if (v == "true" || (allow01 && v == "1")) return true;
  if (v == "false" || (allow01 && v == "0")) return false;
  throw std::runtime_error("Invalid option: " + key + " must be true or false.");
}

inline double to_hours(double v, const std::string& unit) {  // lgar.cxx:653-658
  if (unit == "[s]" || unit == "[sec]" || unit == "") return v / 3600;
  if (unit == "[min]" || unit == "[minute]") return v / 60;
  if (unit == "[h]" || unit == "[hr]") return v / 1.0;
  return v;
}

// InitFromConfigFile, lgar.cxx:231-1016 (same keys, same required-key errors, same unit handling)
inline ParsedConfig parse_config(const std::string& config_file) {
  std::ifstream fp(config_file);
  if (!fp) throw std::runtime_error("cannot open configuration file '" + config_file + "'");
  ParsedConfig c;
  bool s_thick = false, s_psi = false, s_dt = false, s_end = false, s_fr = false, s_soil = false, s_wp = false, s_fc = false;
  bool s_a = false, s_b = false, s_frac = false, s_as = false, s_bs = false, s_fs = false, s_file = false, s_giuh = false;
  std::string line;
  while (std::getline(fp, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    size_t eq = line.find("=");
    int loc_eq = (int)(eq == std::string::npos ? -1 : (int)eq) + 1;
    int loc_u = (int)(line.find("[") == std::string::npos ? -1 : (int)line.find("["));
    std::string key = line.substr(0, eq);
    std::string unit = "";
    if (loc_u >= 0) unit = line.substr(loc_u, line.find("]") + 1);  // (pos, len) exactly as lgar.cxx:323
    std::string val = (loc_u >= 0) ? line.substr(loc_eq, loc_u - loc_eq) : line.substr(loc_eq);
    try {  // std::stod on a malformed number: name the key instead of reporting "stod"
    if (key == "layer_thickness") { c.layer_thickness = read_vector(val); s_thick = true; }
    else if (key == "layer_soil_type") { for (double v : read_vector(val)) c.layer_soil_type.push_back((int)v); s_soil = true; }
    else if (key == "giuh_ordinates") { c.giuh_hourly = read_vector(val); s_giuh = true; }
    else if (key == "initial_psi") { c.initial_psi = std::stod(val); s_psi = true; }
    else if (key == "max_valid_soil_types") {
      c.num_soil_types = std::min(std::stoi(val), MAX_NUM_SOIL_TYPES);
      if (c.num_soil_types < 1) throw std::runtime_error("max_valid_soil_types must be at least 1");  // the soil table has one row per type, 1-based
    }
    else if (key == "soil_params_file") { c.soil_params_file = val; s_file = true; }
    else if (key == "wilting_point_psi") { c.wilting_point_psi = std::stod(val); s_wp = true; }
    else if (key == "field_capacity_psi") { c.field_capacity_psi = std::stod(val); s_fc = true; }
    else if (key == "a") { c.a = std::stod(val); s_a = true; }
    else if (key == "b") { c.b = std::stod(val); s_b = true; }
    else if (key == "frac_to_CR" || key == "frac_to_GW") { c.frac_to_CR = std::stod(val); s_frac = true; }
    else if (key == "a_slow") { c.a_slow = std::stod(val); s_as = true; }
    else if (key == "b_slow") { c.b_slow = std::stod(val); s_bs = true; }
    else if (key == "frac_slow") { c.frac_slow = std::stod(val); s_fs = true; }
    else if (key == "spf_factor") { c.spf_factor = std::stod(val); }
    else if (key == "use_closed_form_G") { c.use_closed_form_G = parse_bool(key, val, false); }
    else if (key == "free_drainage_enabled") { c.free_drainage_enabled = parse_bool(key, val, false); }
    else if (key == "free_drainage_to_CR") { c.free_drainage_to_CR = parse_bool(key, val, false); }
    else if (key == "PET_affects_precip") { c.PET_affects_precip = parse_bool(key, val, false); }
    else if (key == "allow_flux_caching") { c.allow_flux_caching = parse_bool(key, val, false); }
    else if (key == "log_mode") { c.log_mode = parse_bool(key, val, true); }
    else if (key == "mbal_tol") { c.mbal_tol = std::stod(val); }
    else if (key == "adaptive_timestep") { c.adaptive_timestep = parse_bool(key, val, true); }
    else if (key == "timestep") { c.timestep_h = to_hours(std::stod(val), unit); s_dt = true; }
    else if (key == "endtime") {
      double v = std::stod(val);
      if (unit == "[s]" || unit == "[sec]" || unit == "") c.endtime_s = v;
      else if (unit == "[min]" || unit == "[minute]") c.endtime_s = v * 60.0;
      else if (unit == "[h]" || unit == "[hr]") c.endtime_s = v * 3600.0;
      else if (unit == "[d]" || unit == "[day]") c.endtime_s = v * 86400.0;
      s_end = true;
    }
    else if (key == "forcing_resolution") { c.forcing_resolution_h = to_hours(std::stod(val), unit); s_fr = true; }
    else if (key == "sft_coupled") { c.sft_coupled = parse_bool(key, val, false); }
    else if (key == "soil_z") { c.soil_z = read_vector(val); }  // lgar.cxx:390-400
    else if (key == "ponded_depth_max") { c.ponded_depth_max = fmax(std::stod(val), 0.0); }
    else if (key == "calib_params") { c.calib_params = parse_bool(key, val, false); }
    } catch (const std::invalid_argument&) {
      throw std::runtime_error("The configuration file '" + config_file + "': cannot read the value '" + val + "' of " + key + " as a number.");
    } catch (const std::out_of_range&) {
      throw std::runtime_error("The configuration file '" + config_file + "': the value '" + val + "' of " + key + " is out of range.");
    }
  }
  auto need = [&](bool ok, const char* what) {
    if (!ok) throw std::runtime_error("The configuration file '" + config_file + "' does not set " + what + ".");
  };
  need(s_soil, "layer_soil_type");
  need(s_file, "soil_params_file");
  need(s_thick, "layer_thickness");
  need(s_psi, "initial_psi");
  need(s_dt, "timestep");
  need(s_end, "endtime");
  need(s_wp, "wilting_point_psi");
  need(s_fc, "field_capacity_psi");
  if (!((s_a == s_b) && (s_frac == s_b)))
    throw std::runtime_error("The configuration file '" + config_file + "' does not correctly set a, b, and frac_to_CR. Either all or none must be set.");
  if (!((s_as == s_bs) && (s_fs == s_bs)))
    throw std::runtime_error("The configuration file '" + config_file + "' does not correctly set a_slow, b_slow, and frac_slow. Either all or none must be set.");
  if (c.free_drainage_to_CR && !c.free_drainage_enabled)
    throw std::runtime_error("The configuration file '" + config_file + "' sets free_drainage_to_CR as true but sets free_drainage_enabled as false.");
  need(s_fr, "forcing_resolution");
  need(s_giuh, "giuh_ordinates");
  if (!(c.timestep_h > 0) || !(c.forcing_resolution_h > 0) || !(c.endtime_s > 0))
    throw std::runtime_error("timestep, forcing_resolution and endtime must be positive");
  if (c.log_mode) {  // lgar.cxx:796-801
    c.a = pow(10.0, c.a);
    if (s_as) c.a_slow = pow(10.0, c.a_slow);
  }
  if (c.sft_coupled && c.soil_z.empty())  // lgar.cxx:946-952
    throw std::runtime_error("The configuration file '" + config_file + "' does not set soil_z.");
  return c;
}

// lgar_read_vG_param_file, lgar.cxx:2674-2758.  sscanf("%s %lf ...") stops the name at the first
// blank, so a row whose quoted name contains a space converts no number and silently inherits the
// previous row's values (quirk Q2).  Returns rows read; table is 1-based.
// The reference opens soil_params_file relative to the process's working directory only.  As a
// convenience the engine also tries the config file's directory and up to two levels above it.
inline FILE* open_relative(const std::string& path, const std::string& config_file) {
  FILE* fp = fopen(path.c_str(), "r");
  if (fp || path.empty() || path[0] == '/') return fp;
  std::string dir = config_file;
  size_t slash = dir.find_last_of('/');
  dir = (slash == std::string::npos) ? "." : dir.substr(0, slash);
  for (int up = 0; up < 3 && !fp; up++) {
    fp = fopen((dir + "/" + path).c_str(), "r");
    dir += "/..";
  }
  return fp;
}

inline int read_soil_file(const std::string& path, int num_soil_types, double wilting_point_psi_cm, bool log_mode, std::vector<SoilRow>& table,
                          const std::string& config_file = "") {
  FILE* fp = open_relative(path, config_file);
  if (!fp) throw std::runtime_error("Can't open input file named " + path);
  if (num_soil_types < 1) {
    fclose(fp);
    throw std::runtime_error("max_valid_soil_types must be at least 1");
  }
  table.assign(num_soil_types + 1, SoilRow());
  char jstr[256], soil_name[256];
  double theta_e = 0, theta_r = 0, vg_n = 0, vg_alpha = 0, Ksat = 0;
  int nread = 0, soil = 1;
  if (!fgets(jstr, 255, fp)) { fclose(fp); return 0; }
  while (fgets(jstr, 255, fp) != NULL) {
    if (soil > num_soil_types) break;  // never write past the table, whatever the caller passed
    sscanf(jstr, "%s %lf %lf %lf %lf %lf", soil_name, &theta_r, &theta_e, &vg_alpha, &vg_n, &Ksat);
    SoilRow& s = table[soil];
    if (soil > 1) { s.lambda = table[soil - 1].lambda; s.psib = table[soil - 1].psib; }
    double alpha = vg_alpha, ks = Ksat;
    if (log_mode) { alpha = pow(10.0, alpha); ks = pow(10.0, ks); }
    s.theta_r = theta_r; s.theta_e = theta_e; s.alpha = alpha; s.n = vg_n; s.m = 1 - 1 / vg_n; s.Ksat = ks;
    s.theta_wp = theta_from_h(wilting_point_psi_cm, s);
    if (1.0 < vg_n) {
      double m = 1.0 - 1.0 / vg_n;
      double p = 1.0 + 2.0 / m;
      s.lambda = 2.0 / (p - 3.0);
      s.psib = (p + 3.0) * (147.8 + 8.1 * p + 0.092 * p * p) / (2.0 * alpha * p * (p - 1.0) * (55.6 + 7.4 * p + p * p));
    }
    s.h_min = s.psib * (2.0 + 3.0 / s.lambda) / (1.0 + 3.0 / s.lambda);
    nread++;
    soil++;
    if (num_soil_types == nread) break;
  }
  fclose(fp);
  return nread;
}

// engine-wide constants, incl. GIUH re-sampling (lgar.cxx:918-929, quirk Q12)
inline LgarConfig make_engine_config(const ParsedConfig& pc) {
  LgarConfig g;
  memset(&g, 0, sizeof(g));
  g.use_closed_form_G = pc.use_closed_form_G; g.adaptive_timestep = pc.adaptive_timestep; g.allow_flux_caching = pc.allow_flux_caching;
  g.free_drainage_enabled = pc.free_drainage_enabled; g.free_drainage_to_CR = pc.free_drainage_to_CR;
  g.PET_affects_precip = pc.PET_affects_precip;
  g.nint = 120;  // lgar.cxx:992
  g.mbal_tol = pc.mbal_tol;
  g.timestep_h = pc.timestep_h; g.forcing_resolution_h = pc.forcing_resolution_h;
  g.initial_psi_cm = pc.initial_psi; g.wilting_point_psi_cm = pc.wilting_point_psi; g.field_capacity_psi_cm = pc.field_capacity_psi;
  int factor = int(1.0 / pc.forcing_resolution_h);
  int n = factor * (int)pc.giuh_hourly.size();
  if (n < 1 || n > LGAR_MAXG) throw std::runtime_error("number of (re-sampled) GIUH ordinates must be 1.." + std::to_string(LGAR_MAXG));
  g.num_giuh = n;
  for (size_t i = 0; i < pc.giuh_hourly.size(); i++)
    for (int j = 0; j < factor; j++) g.giuh[j + i * factor] = pc.giuh_hourly[i] / double(factor);
  return g;
}

// per-layer part of a class row: the vG parameters and every quantity the reference recomputes per call, from the
// soil_properties_ rows of the layers (1-based); c.num_layers and c.cum are set by the caller
inline void fill_class_layers(LgarColumnClass& c, const ParsedConfig& pc, const SoilRow* rows) {
  const int L = c.num_layers;
  for (int k = 1; k <= L; k++) {
    const SoilRow& s = rows[k];
    LgarLayer& l = c.lay[k];
    l.theta_r = s.theta_r; l.theta_e = s.theta_e; l.alpha = s.alpha; l.n = s.n; l.m = s.m; l.Ksat = s.Ksat;
    l.de = s.theta_e - s.theta_r;
    l.inv_m = 1.0 / s.m;
    l.neg_inv_m = -1.0 / s.m;
    l.inv_n = 1.0 / s.n;
    l.inv_alpha = 1.0 / s.alpha;
    l.H_c = s.psib * ((2 + 3 * s.lambda) / (1 + 3 * s.lambda));
    l.p_exp = (3 + 1 / s.lambda);
    l.theta_min = theta_from_h(1.E7, s);
    l.thick = c.cum[k] - c.cum[k - 1];
    l.h_min = s.h_min;
    l.ln_alpha = log(s.alpha);
    {  // max |d theta / d h| = de m n t (1+t)^(-m-1) / h at t = (alpha h)^n = m
      const double hs = pow(s.m, 1.0 / s.n) / s.alpha;
      const double v = (s.theta_e - s.theta_r) * s.m * s.n * s.m * pow(1.0 + s.m, -s.m - 1.0) / hs;
      l.slope_max = (v == v && v > 0.0) ? v : 1e300;
    }
    c.min_storage += s.theta_r * (c.cum[k] - c.cum[k - 1]);
    c.max_storage += s.theta_e * (c.cum[k] - c.cum[k - 1]);
    c.largest_Ks = fmax(s.Ksat, c.largest_Ks);
    // calc_aet's psi at which AET = PET/2, aet.cxx:47-57
    double theta_fc = theta_from_h(pc.field_capacity_psi, s);
    double wp_head_theta = theta_from_h(pc.wilting_point_psi, s);
    double theta_50 = (theta_fc - wp_head_theta) * 1 / 2 + wp_head_theta;
    double Se = (theta_50 - s.theta_r) / (s.theta_e - s.theta_r);
    c.psi_50_cm[k] = h_from_Se(Se, s.alpha, s.m, s.n);
    // InitializeWettingFronts, lgar.cxx:1035-1061
    c.initial_theta[k] = theta_from_h(pc.initial_psi, s);
    double Se0 = (c.initial_theta[k] - s.theta_r) / (s.theta_e - s.theta_r);
    c.initial_K[k] = K_from_Se(Se0, s.Ksat, s.m);
  }
}

// What do_initialize (lgar_engine.cu) requires of a column's soil types and layer thicknesses before it builds a class row;
// shared with the host build of the kernel text (tests/hostsim) so that both reject the same inputs.
inline void check_column_class(const ParsedConfig& pc, int soils_in_file, int L, const int* types, const double* thick) {
  for (int k = 0; k < L; k++) {
    if (types[k] > soils_in_file && types[k] <= pc.num_soil_types)  // lgar.cxx:814-820
      throw std::runtime_error("Soil type is less than max_valid_soil_types but is greater than the number of lines in the soil data file.");
    if (!(thick[k] > 0)) throw std::runtime_error("layer thickness must be positive");
  }
}

// reservoir / preferential-flow / ponding parameters of a class row from the (possibly calibrated) configuration values
inline void fill_class_params(LgarColumnClass& c, const ParsedConfig& pc) {
  c.par.a = pc.a; c.par.b = pc.b; c.par.frac_to_CR = pc.frac_to_CR;
  c.par.a_slow = pc.a_slow; c.par.b_slow = pc.b_slow; c.par.frac_slow = pc.frac_slow;
  c.par.spf_factor = pc.spf_factor; c.par.ponded_depth_max_cm = pc.ponded_depth_max;
}

inline LgarColumnClass make_class(const ParsedConfig& pc, const std::vector<SoilRow>& soils, int num_layers, const int* types, const double* thick) {
  LgarColumnClass c;
  memset(&c, 0, sizeof(c));
  fill_class_params(c, pc);
  const int L = num_layers;
  c.num_layers = L;
  c.cum[0] = 0.0;
  for (int k = 1; k <= L; k++) c.cum[k] = c.cum[k - 1] + thick[k - 1];  // lgar.cxx:339-342
  c.invalid_soil_type = 0;
  for (int k = 1; k <= L; k++) {
    if (types[k - 1] > pc.num_soil_types) {  // lgar.cxx:824-836
      c.invalid_soil_type = 1;
      break;
    }
    if (types[k - 1] < 1) throw std::runtime_error("layer_soil_type must be >= 1");
  }
  if (c.invalid_soil_type) return c;
  std::vector<SoilRow> rows((size_t)L + 1);
  for (int k = 1; k <= L; k++) rows[k] = soils[types[k - 1]];
  fill_class_layers(c, pc, rows.data());
  return c;
}

// the same for explicit per-layer rows (1-based): calibrated parameters, src/bmi_lgar.cxx:916-1078
inline LgarColumnClass make_class_from_rows(const ParsedConfig& pc, const SoilRow* rows, int num_layers, const double* thick) {
  LgarColumnClass c;
  memset(&c, 0, sizeof(c));
  fill_class_params(c, pc);
  c.num_layers = num_layers;
  c.cum[0] = 0.0;
  for (int k = 1; k <= num_layers; k++) c.cum[k] = c.cum[k - 1] + thick[k - 1];
  fill_class_layers(c, pc, rows);
  return c;
}

inline double class_initial_storage(const LgarColumnClass& c) {  // lgar_calc_mass_bal on the initial list
  double sum = 0.0;
  for (int k = 1; k <= c.num_layers; k++) {
    double base = c.cum[k - 1];
    if (k < c.num_layers)
      sum += (c.cum[k] - base) * c.initial_theta[k];
    else
      sum += c.initial_theta[k] * (c.cum[k] - base);
  }
  return sum;
}

}  // namespace lgarhost
#endif
