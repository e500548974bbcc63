// lgar_c_b200/csrc/lasam_b200_standalone.cpp -- batched standalone driver over the C ABI (include/lgarb.h).
//
// The caller the reference ships is src/bmi_main_lgar.cxx (`lasam_standalone CONFIG`): it reads the forcing CSV
// named in the config (ReadForcingData, :183-262), loops SetValue(P), SetValue(PET), Update(), GetValue(...) over
// int(endtime / timestep) steps (:79-160) and writes `data_variables.csv` (Time + 11 variables, "%6.15f", :100-141)
// and `data_layers.csv` (write_state, :266-283).  This program is that driver for a batch of columns
// (SURVEY.md section 8f, rows N1 and N2): same configuration file, same forcing CSV, same two output files for the
// columns one asks for, plus
//   * N columns in one engine (--columns), each reading the forcing with its own circular time shift (--shift), or
//     a columnar binary forcing file [time][column] in f32 or f64 (--forcing-bin);
//   * the run is streamed through lgarb_update_steps_host_stream in chunks (--chunk): host-to-device forcing copies
//     and device-to-host output copies overlap the computation;
//   * a columnar binary output file for every column (--out-bin, --out-vars, --f32).
// With --layers it steps hour by hour like the reference (the front list of every step is needed for
// data_layers.csv).  Host code only: everything numerical happens behind the C ABI on the GPU.
//
// Binary layout (forcing and output, little endian):
//   char magic[4] = "LGF1"; int32 dtype (0 = f64, 1 = f32); int64 n_steps; int64 n_columns; int32 n_vars; int32 pad;
//   char names[n_vars][64]; then n_vars planes [n_steps][n_columns].  A forcing file has the planes
//   precipitation_rate and potential_evapotranspiration_rate (mm/h).
#include <cerrno>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/lgarb.h"
#include "../../include/lgarb_group.h"

namespace {

struct BinHeader {
  char magic[4];
  int32_t dtype;
  int64_t n_steps;
  int64_t n_columns;
  int32_t n_vars;
  int32_t pad;
};

const char* kVarNames[11] = {"precipitation", "potential_evapotranspiration", "actual_evapotranspiration", "surface_runoff", "giuh_runoff",
                             "soil_storage", "total_discharge", "infiltration", "percolation", "conceptual_reservoir_to_stream_discharge",
                             "mass_balance"};  // bmi_main_lgar.cxx:62-72

[[noreturn]] void die(const std::string& msg) {
  fprintf(stderr, "lasam_b200_standalone: %s\n", msg.c_str());
  exit(1);
}

void ck(lgarb_engine* e, int rc, const char* what) {
  if (rc != LGARB_SUCCESS) die(std::string(what) + ": " + lgarb_last_error(e));
}

// forcing_file=... of the configuration file (bmi_main_lgar.cxx:196-222)
std::string forcing_file_of(const std::string& config) {
  std::ifstream f(config);
  if (!f) die(config + " does not exist");
  std::string line;
  while (std::getline(f, line)) {
    const size_t eq = line.find('=');
    if (eq != std::string::npos && line.substr(0, eq) == "forcing_file") {
      std::string v = line.substr(eq + 1);
      while (!v.empty() && (v.back() == '\r' || v.back() == ' ')) v.pop_back();
      return v;
    }
  }
  die(config + " does not provide forcing_file");
}

// the reference's CSV reader (bmi_main_lgar.cxx:232-262): header line, then time, P [mm/h], PET [mm/h]
void read_forcing_csv(const std::string& path, std::vector<std::string>& time, std::vector<double>& P, std::vector<double>& E) {
  std::ifstream fp(path);
  if (!fp) die("file " + path + " doesn't exist");
  std::string line, cell;
  std::getline(fp, line);
  while (std::getline(fp, line)) {
    std::stringstream ls(line);
    int count = 0;
    while (std::getline(ls, cell, ',')) {
      if (count == 0) time.push_back(cell);
      else if (count == 1) P.push_back(std::stod(cell));
      else if (count == 2) E.push_back(std::stod(cell));
      count++;
    }
  }
}

std::vector<std::string> split(const std::string& s, char sep) {
  std::vector<std::string> out;
  std::stringstream ss(s);
  std::string item;
  while (std::getline(ss, item, sep))
    if (!item.empty()) out.push_back(item);
  return out;
}

struct Options {
  std::string config, out_dir = ".", forcing_bin, out_bin;
  int64_t columns = 1, shift = 0, chunk = 240;
  int device = 0;
  std::vector<int> devices;  // --devices a,b,..: the columns in contiguous blocks over these GPUs (include/lgarb_group.h)
  bool layers = false, f32 = false, quiet = false;
  std::vector<int64_t> write_columns = {0};
  std::vector<std::string> out_vars = {"total_discharge"};
};

Options parse(int argc, char** argv) {
  Options o;
  for (int i = 1; i < argc; i++) {
    const std::string a = argv[i];
    auto need = [&](const char* name) -> std::string {
      if (i + 1 >= argc) die(std::string("missing value after ") + name);
      return argv[++i];
    };
    if (a == "--columns") o.columns = std::stoll(need("--columns"));
    else if (a == "--shift") o.shift = std::stoll(need("--shift"));
    else if (a == "--chunk") o.chunk = std::stoll(need("--chunk"));
    else if (a == "--device") o.device = std::stoi(need("--device"));
    else if (a == "--devices") {
      for (auto& d : split(need("--devices"), ',')) o.devices.push_back(std::stoi(d));
    }
    else if (a == "--out-dir") o.out_dir = need("--out-dir");
    else if (a == "--forcing-bin") o.forcing_bin = need("--forcing-bin");
    else if (a == "--out-bin") o.out_bin = need("--out-bin");
    else if (a == "--out-vars") o.out_vars = split(need("--out-vars"), ',');
    else if (a == "--write-columns") {
      const std::string v = need("--write-columns");
      o.write_columns.clear();
      if (v != "none")
        for (auto& s : split(v, ',')) o.write_columns.push_back(std::stoll(s));
    } else if (a == "--layers") o.layers = true;
    else if (a == "--f32") o.f32 = true;
    else if (a == "--quiet") o.quiet = true;
    else if (!a.empty() && a[0] == '-') die("unknown option " + a);
    else if (o.config.empty()) o.config = a;
    else die("more than one configuration file");
  }
  return o;
}

std::string suffix_for(const Options& o, int64_t col) { return o.columns == 1 ? std::string() : "_c" + std::to_string(col); }

// --devices: the batch in contiguous column blocks over several GPUs of the node (liblgarb_group.so).  Forcing from the CSV of the
// configuration file with a circular shift per column, data_variables CSVs for the columns asked for, and at the end the balance
// terms of every column gathered onto the first device over NCCL with their totals.
int run_group(const Options& o) {
  if (o.layers || o.f32 || !o.forcing_bin.empty() || !o.out_bin.empty()) die("--devices runs the CSV forcing with f64 CSV outputs (no --layers, --f32, --forcing-bin, --out-bin)");
  lgarb_group* g = nullptr;
  if (lgarb_group_create(&g, (int)o.devices.size(), o.devices.data()) != LGARB_SUCCESS) die("lgarb_group_create failed (device list)");
  auto gck = [&](int rc, const char* what) {
    if (rc != LGARB_SUCCESS) die(std::string(what) + ": " + lgarb_group_last_error(g));
  };
  gck(lgarb_group_initialize(g, o.config.c_str(), o.columns, nullptr, nullptr), "Initialize");
  lgarb_engine* e0 = lgarb_group_engine(g, 0);
  double endtime = 0, timestep = 0;
  ck(e0, lgarb_get_end_time(e0, &endtime), "GetEndTime");
  ck(e0, lgarb_get_time_step(e0, &timestep), "GetTimeStep");
  const int64_t nsteps = (int64_t)(endtime / timestep), N = o.columns;
  std::vector<std::string> time;
  std::vector<double> P1, E1;
  read_forcing_csv(forcing_file_of(o.config), time, P1, E1);
  if (nsteps > (int64_t)E1.size()) die("nsteps exceeds the forcing data");
  const int64_t nf = (int64_t)E1.size();
  for (int64_t c : o.write_columns)
    if (c < 0 || c >= N) die("--write-columns: column out of range");
  std::vector<FILE*> fvar;
  for (int64_t c : o.write_columns) {
    FILE* fv = fopen((o.out_dir + "/data_variables" + suffix_for(o, c) + ".csv").c_str(), "w");
    if (!fv) die("cannot write to " + o.out_dir);
    fprintf(fv, "Time,");
    for (int j = 0; j < 11; j++) fprintf(fv, "%s%s", kVarNames[j], j == 10 ? "\n" : ",");
    fvar.push_back(fv);
  }
  const int nout = o.write_columns.empty() ? 0 : 11;
  const int64_t seg = std::max<int64_t>(1, std::min<int64_t>(nsteps, (int64_t)((4ull << 30) / ((size_t)(2 + nout) * (size_t)N * 8))));
  std::vector<double> P((size_t)seg * N), E((size_t)seg * N);
  std::vector<std::vector<double>> outs(nout, std::vector<double>((size_t)seg * N));
  std::vector<double*> ptrs;
  for (auto& v : outs) ptrs.push_back(v.data());
  for (int64_t t0 = 0; t0 < nsteps; t0 += seg) {
    const int64_t len = std::min<int64_t>(seg, nsteps - t0);
    for (int64_t t = 0; t < len; t++)
      for (int64_t c = 0; c < N; c++) {
        const int64_t src = (t0 + t + c * o.shift) % nf;
        P[t * N + c] = P1[src];
        E[t * N + c] = E1[src];
      }
    gck(lgarb_group_update_steps_host(g, len, P.data(), E.data(), nout, kVarNames, ptrs.data()), "update_steps_host");
    for (size_t w = 0; w < o.write_columns.size(); w++) {
      const int64_t c = o.write_columns[w];
      for (int64_t t = 0; t < len; t++) {
        fprintf(fvar[w], "%s,", time[(t0 + t + c * o.shift) % nf].c_str());
        for (int j = 0; j < 11; j++) fprintf(fvar[w], "%6.15f%s", outs[j][t * N + c], j == 10 ? "\n" : ",");
      }
    }
  }
  for (FILE* f : fvar) fclose(f);
  std::vector<double> gm((size_t)N * 16);
  double tot[16], ms = 0.0;
  gck(lgarb_group_gather_mass_balance(g, gm.data(), tot, &ms), "gather");
  if (!o.quiet) {
    double worst = 0.0;
    for (int64_t c = 0; c < N; c++)
      if (std::isfinite(gm[c * 16 + 15])) worst = std::fmax(worst, std::fabs(gm[c * 16 + 15]));
    printf("columns %lld over %d devices, steps %lld, max |global mass balance| %.6e cm; totals over all columns: precipitation %.6f cm, infiltration %.6f cm, "
           "AET %.6f cm, discharge %.6f cm (NCCL gather of %lld x 16 terms + reduce: %.3f ms on the first device)\n",
           (long long)N, lgarb_group_num_devices(g), (long long)nsteps, worst, tot[1], tot[2], tot[8], tot[12], (long long)N, ms);
    for (int64_t c : o.write_columns)
      printf("column %lld: initial water %.10f cm, precipitation %.10f cm, infiltration %.10f cm, runoff %.10f cm, AET %.10f cm, final water %.10f cm, "
             "global balance %.6e cm\n",
             (long long)c, gm[c * 16 + 0], gm[c * 16 + 1], gm[c * 16 + 2], gm[c * 16 + 5], gm[c * 16 + 8], gm[c * 16 + 3], gm[c * 16 + 15]);
  }
  lgarb_group_destroy(g);
  return 0;
}

}  // namespace

int main(int argc, char** argv) {
  Options o = parse(argc, argv);
  if (o.config.empty()) {
    printf("Usage: lasam_b200_standalone CONFIGURATION_FILE [--columns N] [--shift S] [--chunk H] [--layers]\n"
           "         [--write-columns a,b,..|none] [--out-dir DIR] [--forcing-bin FILE] [--out-bin FILE] [--out-vars v1,v2] [--f32] [--devices a,b,..]\n"
           "Runs LASAM for a batch of columns on the GPU through its batched BMI (include/lgarb.h).\n"
           "Outputs are written to `data_variables.csv` (and `data_layers.csv` with --layers) for the columns asked for.\n");
    return 0;
  }
  if (o.devices.size() > 1) return run_group(o);
  if (o.devices.size() == 1) o.device = o.devices[0];
  lgarb_engine* e = nullptr;
  if (lgarb_create(&e) != LGARB_SUCCESS || !e) die("lgarb_create failed");
  ck(e, lgarb_initialize(e, o.config.c_str(), o.columns, o.device), "Initialize");
  double endtime = 0, timestep = 0;
  ck(e, lgarb_get_end_time(e, &endtime), "GetEndTime");
  ck(e, lgarb_get_time_step(e, &timestep), "GetTimeStep");
  const int64_t nsteps = (int64_t)(endtime / timestep);  // bmi_main_lgar.cxx:81
  const int64_t N = o.columns;

  // ---- forcing ----
  std::vector<std::string> time;
  std::vector<double> P1, E1;       // the CSV series (one column)
  std::vector<double> P, E;         // [nsteps][N] f64
  std::vector<float> P32, E32;      // or f32
  bool in32 = false;
  if (!o.forcing_bin.empty()) {
    FILE* f = fopen(o.forcing_bin.c_str(), "rb");
    if (!f) die("cannot open " + o.forcing_bin);
    BinHeader h;
    if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, "LGF1", 4) != 0) die(o.forcing_bin + ": not an LGF1 file");
    if (h.n_columns != N || h.n_steps < nsteps || h.n_vars != 2) die(o.forcing_bin + ": needs 2 planes of at least " + std::to_string(nsteps) + " x " + std::to_string(N));
    fseek(f, (long)(h.n_vars * 64), SEEK_CUR);
    in32 = h.dtype == 1;
    const size_t plane = (size_t)h.n_steps * (size_t)N, want = (size_t)nsteps * (size_t)N;
    if (in32) {
      P32.resize(want); E32.resize(want);
      if (fread(P32.data(), 4, want, f) != want) die("short read");
      fseek(f, (long)((plane - want) * 4), SEEK_CUR);
      if (fread(E32.data(), 4, want, f) != want) die("short read");
    } else {
      P.resize(want); E.resize(want);
      if (fread(P.data(), 8, want, f) != want) die("short read");
      fseek(f, (long)((plane - want) * 8), SEEK_CUR);
      if (fread(E.data(), 8, want, f) != want) die("short read");
    }
    fclose(f);
    for (int64_t t = 0; t < nsteps; t++) time.push_back(std::to_string(t));
  } else {
    read_forcing_csv(forcing_file_of(o.config), time, P1, E1);
    if (nsteps > (int64_t)E1.size()) die("nsteps exceeds the forcing data");  // assert at bmi_main_lgar.cxx:90
    const int64_t nf = (int64_t)E1.size();
    if (!(o.layers)) {
      if (o.f32) {
        in32 = true;
        P32.resize((size_t)nsteps * N); E32.resize((size_t)nsteps * N);
      } else {
        P.resize((size_t)nsteps * N); E.resize((size_t)nsteps * N);
      }
      for (int64_t t = 0; t < nsteps; t++)
        for (int64_t c = 0; c < N; c++) {
          const int64_t src = (t + c * o.shift) % nf;
          if (in32) { P32[t * N + c] = (float)P1[src]; E32[t * N + c] = (float)E1[src]; }
          else { P[t * N + c] = P1[src]; E[t * N + c] = E1[src]; }
        }
    }
  }
  for (int64_t c : o.write_columns)
    if (c < 0 || c >= N) die("--write-columns: column out of range");

  // ---- output files of the reference for the selected columns ----
  std::vector<FILE*> fvar, flay;
  for (int64_t c : o.write_columns) {
    const std::string sfx = suffix_for(o, c);
    FILE* fv = fopen((o.out_dir + "/data_variables" + sfx + ".csv").c_str(), "w");
    if (!fv) die("cannot write to " + o.out_dir);
    fprintf(fv, "Time,");
    for (int j = 0; j < 11; j++) fprintf(fv, "%s%s", kVarNames[j], j == 10 ? "\n" : ",");
    fvar.push_back(fv);
    if (o.layers) flay.push_back(fopen((o.out_dir + "/data_layers" + sfx + ".csv").c_str(), "w"));
  }

  if (o.layers) {
    // the reference's loop, one forcing step at a time (bmi_main_lgar.cxx:116-160)
    if (!o.forcing_bin.empty()) die("--layers reads the forcing CSV of the configuration file");
    const int64_t nf = (int64_t)E1.size();
    std::vector<double> p(N), pe(N), val(N);
    std::vector<double> depth((size_t)N * LGARB_MAX_FRONTS), theta((size_t)N * LGARB_MAX_FRONTS), psi((size_t)N * LGARB_MAX_FRONTS);
    std::vector<int32_t> layer((size_t)N * LGARB_MAX_FRONTS), nfr(N);
    for (int64_t t = 0; t < nsteps; t++) {
      for (int64_t c = 0; c < N; c++) {
        const int64_t src = (t + c * o.shift) % nf;
        p[c] = P1[src];
        pe[c] = E1[src];
      }
      ck(e, lgarb_set_value(e, "precipitation_rate", p.data()), "SetValue");
      ck(e, lgarb_set_value(e, "potential_evapotranspiration_rate", pe.data()), "SetValue");
      ck(e, lgarb_update(e), "Update");
      std::vector<std::vector<double>> row(o.write_columns.size(), std::vector<double>(11));
      for (int j = 0; j < 11; j++) {
        ck(e, lgarb_get_value(e, kVarNames[j], val.data()), "GetValue");
        for (size_t w = 0; w < o.write_columns.size(); w++) row[w][j] = val[o.write_columns[w]];
      }
      ck(e, lgarb_get_fronts(e, 0, N, depth.data(), theta.data(), psi.data(), nullptr, nullptr, layer.data(), nullptr, nfr.data()), "get_fronts");
      for (size_t w = 0; w < o.write_columns.size(); w++) {
        const int64_t c = o.write_columns[w];
        fprintf(fvar[w], "%s,", time[(t + c * o.shift) % nf].c_str());
        for (int j = 0; j < 11; j++) fprintf(fvar[w], "%6.15f%s", row[w][j], j == 10 ? "\n" : ",");
        fprintf(flay[w], "# Timestep = %d, %s \n", (int)t, time[(t + c * o.shift) % nf].c_str());
        fprintf(flay[w], "[");
        for (int i = 0; i < nfr[c]; i++)  // write_state, bmi_main_lgar.cxx:266-283 (depth and psi in mm)
          fprintf(flay[w], "%s(%lf,%lf,%d,%d,%lf)", i ? "|" : "", depth[c * LGARB_MAX_FRONTS + i] * 10., theta[c * LGARB_MAX_FRONTS + i], layer[c * LGARB_MAX_FRONTS + i], i + 1, psi[c * LGARB_MAX_FRONTS + i] * 10.);
        fprintf(flay[w], "]\n");
      }
    }
  } else {
    // streamed run: chunks of o.chunk hours, copies overlapped with compute
    const bool want_csv = !o.write_columns.empty();
    std::vector<std::string> names;
    if (want_csv)
      for (int j = 0; j < 11; j++) names.push_back(kVarNames[j]);
    std::vector<int> bin_idx;  // position of each --out-vars variable in `names`
    if (!o.out_bin.empty())
      for (auto& v : o.out_vars) {
        size_t k = 0;
        while (k < names.size() && names[k] != v) k++;
        if (k == names.size()) names.push_back(v);
        bin_idx.push_back((int)k);
      }
    // keep one segment of outputs below ~4 GB of host memory
    const size_t item = o.f32 ? 4 : 8;
    int64_t seg = nsteps;
    if (!names.empty()) seg = std::max<int64_t>(1, std::min<int64_t>(nsteps, (int64_t)((4ull << 30) / (names.size() * (size_t)N * item))));
    std::vector<std::vector<char>> outs(names.size());
    for (auto& v : outs) v.resize((size_t)seg * (size_t)N * item);
    std::vector<const char*> cnames;
    for (auto& n : names) cnames.push_back(n.c_str());
    FILE* fb = nullptr;
    std::vector<long> plane_off;
    if (!o.out_bin.empty()) {
      fb = fopen(o.out_bin.c_str(), "wb");
      if (!fb) die("cannot write " + o.out_bin);
      BinHeader h;
      memcpy(h.magic, "LGF1", 4);
      h.dtype = o.f32 ? 1 : 0;
      h.n_steps = nsteps;
      h.n_columns = N;
      h.n_vars = (int32_t)o.out_vars.size();
      h.pad = 0;
      fwrite(&h, sizeof(h), 1, fb);
      for (auto& v : o.out_vars) {
        char nm[64] = {0};
        strncpy(nm, v.c_str(), 63);
        fwrite(nm, 64, 1, fb);
      }
      for (size_t k = 0; k < o.out_vars.size(); k++) plane_off.push_back((long)(sizeof(h) + 64 * o.out_vars.size() + k * (size_t)nsteps * (size_t)N * item));
    }
    const int64_t nf = (int64_t)E1.size();
    for (int64_t t0 = 0; t0 < nsteps; t0 += seg) {
      const int64_t len = std::min<int64_t>(seg, nsteps - t0);
      std::vector<void*> ptrs;
      for (auto& v : outs) ptrs.push_back(v.data());
      const void* hp = in32 ? (const void*)(P32.data() + (size_t)t0 * N) : (const void*)(P.data() + (size_t)t0 * N);
      const void* he = in32 ? (const void*)(E32.data() + (size_t)t0 * N) : (const void*)(E.data() + (size_t)t0 * N);
      ck(e, lgarb_update_steps_host_stream(e, len, o.chunk, in32 ? LGARB_F32 : LGARB_F64, hp, he, (int)names.size(), cnames.data(),
                                           o.f32 ? LGARB_F32 : LGARB_F64, ptrs.data()),
         "update_steps_host_stream");
      auto at = [&](size_t k, int64_t t, int64_t c) -> double {
        return o.f32 ? (double)((const float*)outs[k].data())[t * N + c] : ((const double*)outs[k].data())[t * N + c];
      };
      for (size_t w = 0; w < o.write_columns.size(); w++) {
        const int64_t c = o.write_columns[w];
        for (int64_t t = 0; t < len; t++) {
          const std::string& stamp = o.forcing_bin.empty() ? time[(t0 + t + c * o.shift) % nf] : time[t0 + t];
          fprintf(fvar[w], "%s,", stamp.c_str());
          for (int j = 0; j < 11; j++) fprintf(fvar[w], "%6.15f%s", at(j, t, c), j == 10 ? "\n" : ",");
        }
      }
      if (fb)
        for (size_t k = 0; k < o.out_vars.size(); k++) {
          fseek(fb, plane_off[k] + (long)((size_t)t0 * (size_t)N * item), SEEK_SET);
          fwrite(outs[bin_idx[k]].data(), item, (size_t)len * (size_t)N, fb);
        }
    }
    if (fb) fclose(fb);
  }
  for (FILE* f : fvar) fclose(f);
  for (FILE* f : flay) fclose(f);

  // global mass balance of the written columns, as Finalize() prints it (lgar_global_mass_balance, lgar.cxx:1185-1214)
  if (!o.quiet) {
    std::vector<double> gm((size_t)N * 16);
    ck(e, lgarb_get_global_mass_balance(e, 0, N, gm.data()), "global mass balance");
    std::vector<int32_t> status(N);
    ck(e, lgarb_get_status(e, 0, N, status.data()), "status");
    int64_t failed = 0;
    double worst = 0.0;
    for (int64_t c = 0; c < N; c++) {
      if (status[c] & LGARB_STATUS_FAIL_MASK) failed++;
      else if (std::isfinite(gm[c * 16 + 15])) worst = std::fmax(worst, std::fabs(gm[c * 16 + 15]));
    }
    printf("columns %lld, steps %lld, frozen columns %lld, max |global mass balance| %.6e cm\n", (long long)N, (long long)nsteps, (long long)failed, worst);
    for (int64_t c : o.write_columns)
      printf("column %lld: initial water %.10f cm, precipitation %.10f cm, infiltration %.10f cm, runoff %.10f cm, AET %.10f cm, final water %.10f cm, "
             "global balance %.6e cm\n",
             (long long)c, gm[c * 16 + 0], gm[c * 16 + 1], gm[c * 16 + 2], gm[c * 16 + 5], gm[c * 16 + 8], gm[c * 16 + 3], gm[c * 16 + 15]);
  }
  lgarb_finalize(e);
  lgarb_destroy(e);
  return 0;
}
