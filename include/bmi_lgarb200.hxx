// include/bmi_lgarb200.hxx -- header-only C++ BMI adapter over the C ABI (include/lgarb.h).
//
// Replaces the reference's `class BmiLGAR : public bmi::Bmi` (include/bmi_lgar.hxx:30-184) and its two
// C symbols bmi_model_create / bmi_model_destroy (include/bmi_lgar.hxx:187-215) for frameworks that load
// a BMI shared library (ngen: realizations/realization_config_lasam.json).  One object = one engine with
// `n_columns` columns (default 1, i.e. exactly the reference's semantics); include the CSDMS `bmi.hxx`
// (the reference ships it as bmi/bmi.hxx) before this header, or put it on the include path.
//
//   g++ -std=c++14 -shared -fPIC -DLGARB200_DEFINE_BMI_MODEL_CREATE adapter.cpp -I<dir of bmi.hxx> -Iinclude \
//       -Llgar_c_b200 -llgarb -o liblasambmi_b200.so
#ifndef BMI_LGARB200_HXX
#define BMI_LGARB200_HXX

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "bmi.hxx"
#include "lgarb.h"

class BmiLGARB200 : public bmi::Bmi {
 public:
  explicit BmiLGARB200(int64_t n_columns = 1, int device = 0) : n_columns_(n_columns), device_(device) {
    if (lgarb_create(&e_) != LGARB_SUCCESS) throw std::runtime_error("lgarb_create failed");
  }
  ~BmiLGARB200() override {
    if (e_) lgarb_destroy(e_);
  }
  BmiLGARB200(const BmiLGARB200&) = delete;
  BmiLGARB200& operator=(const BmiLGARB200&) = delete;

  // Model control functions (reference src/bmi_lgar.cxx:72,133,898,1080)
  void Initialize(std::string config_file) override { ck(lgarb_initialize(e_, config_file.c_str(), n_columns_, device_)); }
  void Update() override { ck(lgarb_update(e_)); }
  void UpdateUntil(double time) override { ck(lgarb_update_until(e_, time)); }
  void Finalize() override { ck(lgarb_finalize(e_)); }

  // Model information functions (:1493-1535)
  std::string GetComponentName() override { return str([&](char* b, int n) { return lgarb_get_component_name(e_, b, n); }); }
  int GetInputItemCount() override { int n; ck(lgarb_get_input_item_count(e_, &n)); return n; }
  int GetOutputItemCount() override { int n; ck(lgarb_get_output_item_count(e_, &n)); return n; }
  std::vector<std::string> GetInputVarNames() override {
    std::vector<std::string> v;
    for (int i = 0; i < GetInputItemCount(); i++) v.push_back(lgarb_get_input_var_name(e_, i));
    return v;
  }
  std::vector<std::string> GetOutputVarNames() override {
    std::vector<std::string> v;
    for (int i = 0; i < GetOutputItemCount(); i++) v.push_back(lgarb_get_output_var_name(e_, i));
    return v;
  }

  // Variable information functions (:1112-1249)
  int GetVarGrid(std::string name) override { int g; ck(lgarb_get_var_grid(e_, name.c_str(), &g)); return g; }
  std::string GetVarType(std::string name) override { return str([&](char* b, int n) { return lgarb_get_var_type(e_, name.c_str(), b, n); }); }
  std::string GetVarUnits(std::string name) override { return str([&](char* b, int n) { return lgarb_get_var_units(e_, name.c_str(), b, n); }); }
  int GetVarItemsize(std::string name) override { int n; ck(lgarb_get_var_itemsize(e_, name.c_str(), &n)); return n; }
  int GetVarNbytes(std::string name) override {
    // the reference's per-front arrays have the live front count as their length (:1297-1298)
    if (n_columns_ == 1 && is_front_array(name)) return (int)sizeof(double) * num_fronts();
    int64_t n;
    ck(lgarb_get_var_nbytes(e_, name.c_str(), &n));
    return (int)n;
  }
  std::string GetVarLocation(std::string name) override { return str([&](char* b, int n) { return lgarb_get_var_location(e_, name.c_str(), b, n); }); }

  // Time functions (:1538-1565)
  double GetCurrentTime() override { double t; ck(lgarb_get_current_time(e_, &t)); return t; }
  double GetStartTime() override { double t; ck(lgarb_get_start_time(e_, &t)); return t; }
  double GetEndTime() override { double t; ck(lgarb_get_end_time(e_, &t)); return t; }
  std::string GetTimeUnits() override { return str([&](char* b, int n) { return lgarb_get_time_units(e_, b, n); }); }
  double GetTimeStep() override { double t; ck(lgarb_get_time_step(e_, &t)); return t; }

  // Variable getters / setters (:1307-1490)
  void GetValue(std::string name, void* dest) override {
    if (n_columns_ == 1 && is_front_array(name)) {  // un-pad: the reference returns exactly num_wetting_fronts items
      double padded[LGARB_MAX_FRONTS];
      ck(lgarb_get_value(e_, name.c_str(), padded));
      std::memcpy(dest, padded, sizeof(double) * (size_t)num_fronts());
      return;
    }
    ck(lgarb_get_value(e_, name.c_str(), dest));
  }
  void* GetValuePtr(std::string) override { throw std::logic_error("GetValuePtr: the engine's arrays live in device memory; use lgarb_get_value_ptr"); }
  void GetValueAtIndices(std::string name, void* dest, int* inds, int count) override { ck(lgarb_get_value_at_indices(e_, name.c_str(), dest, inds, count)); }
  void SetValue(std::string name, void* src) override { ck(lgarb_set_value(e_, name.c_str(), src)); }
  void SetValueAtIndices(std::string name, int* inds, int count, void* src) override { ck(lgarb_set_value_at_indices(e_, name.c_str(), inds, count, src)); }

  // Grid information functions (:1252-1303, :1567-1664)
  int GetGridRank(const int grid) override { int r; ck(lgarb_get_grid_rank(e_, grid, &r)); return r; }
  int GetGridSize(const int grid) override {
    if (n_columns_ == 1 && grid == 3) return num_fronts();
    int64_t n;
    ck(lgarb_get_grid_size(e_, grid, &n));
    return (int)n;
  }
  std::string GetGridType(const int grid) override { return str([&](char* b, int n) { return lgarb_get_grid_type(e_, grid, b, n); }); }
  void GetGridShape(const int grid, int* shape) override {
    if (grid == 2) shape[0] = lgarb_get_num_layers(e_);
    else if (grid == 3) shape[1] = num_fronts();  // sic: the reference writes shape[1] (:1257-1258)
  }
  void GetGridSpacing(const int, double*) override {}
  void GetGridOrigin(const int, double*) override {}
  void GetGridX(const int, double*) override { throw std::logic_error("Not Implemented Function in LGAR"); }
  void GetGridY(const int, double*) override { throw std::logic_error("Not Implemented Function in LGAR"); }
  void GetGridZ(const int, double*) override { throw std::logic_error("Not Implemented Function in LGAR"); }
  int GetGridNodeCount(const int) override { throw std::logic_error("Not Implemented Function in LGAR"); }
  int GetGridEdgeCount(const int) override { throw std::logic_error("Not Implemented Function in LGAR"); }
  int GetGridFaceCount(const int) override { throw std::logic_error("Not Implemented Function in LGAR"); }
  void GetGridEdgeNodes(const int, int*) override { throw std::logic_error("Not Implemented Function in LGAR"); }
  void GetGridFaceEdges(const int, int*) override { throw std::logic_error("Not Implemented Function in LGAR"); }
  void GetGridFaceNodes(const int, int*) override { throw std::logic_error("Not Implemented Function in LGAR"); }
  void GetGridNodesPerFace(const int, int*) override { throw std::logic_error("Not Implemented Function in LGAR"); }

  lgarb_engine* engine() { return e_; }

 private:
  lgarb_engine* e_ = nullptr;
  int64_t n_columns_;
  int device_;
  void ck(int rc) {
    if (rc != LGARB_SUCCESS) throw std::runtime_error(lgarb_last_error(e_));  // the reference throws std::runtime_error too
  }
  template <typename F>
  std::string str(F f) {
    char buf[256];
    ck(f(buf, 256));
    return buf;
  }
  static bool is_front_array(const std::string& n) { return n == "soil_moisture_wetting_fronts" || n == "soil_depth_wetting_fronts"; }
  int num_fronts() {
    int n = 0;  // column 0 only: lgarb_get_value would copy n_columns int32 values into this one int
    ck(lgarb_get_value_range(e_, "soil_num_wetting_fronts", 0, 1, &n));
    return n;
  }
};

#ifdef LGARB200_DEFINE_BMI_MODEL_CREATE
extern "C" {
bmi::Bmi* bmi_model_create() { return new BmiLGARB200(); }
void bmi_model_destroy(bmi::Bmi* ptr) { delete ptr; }
}
#endif

#endif  // BMI_LGARB200_HXX
