// lgar_c_b200/csrc/lgar_group.cu -- liblgarb_group.so: several GPUs of one node behind one handle (include/lgarb_group.h).
//
// Host code over the C ABI of liblgarb.so plus NCCL for the end-of-run gather.  Contiguous column blocks per device, one engine
// and one host thread per device, no collective on the step path (columns share nothing: the reference keeps one model_state per
// object, include/all.hxx:264-275); the gather moves the 16 terms of lgar_global_mass_balance (reference src/lgar.cxx:1162-1216)
// of every column to the first device and reduces their sums there.
#include <cuda_runtime.h>
#include <nccl.h>

#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/lgarb_group.h"

struct lgarb_group {
  std::vector<int> dev;
  std::vector<lgarb_engine*> eng;
  std::vector<int64_t> col0, ncol;
  int64_t ntot = 0;
  bool initialized = false;
  std::vector<ncclComm_t> comm;      // one communicator per device (single process: ncclCommInitAll)
  std::vector<double*> d_sum;        // per device: [ncol_k][16]
  std::vector<double*> d_tot;        // per device: [16] sums of its columns
  double* d_all = nullptr;           // first device: [ntot][16]
  double* d_tot_all = nullptr;       // first device: [16]
  std::string err;
};

namespace {

#define GROUP_CUDA(x)                                                                                        \
  do {                                                                                                       \
    cudaError_t _e = (x);                                                                                    \
    if (_e != cudaSuccess) throw std::runtime_error(std::string(#x) + ": " + cudaGetErrorString(_e));        \
  } while (0)
#define GROUP_NCCL(x)                                                                                        \
  do {                                                                                                       \
    ncclResult_t _r = (x);                                                                                   \
    if (_r != ncclSuccess) throw std::runtime_error(std::string(#x) + ": " + ncclGetErrorString(_r));        \
  } while (0)

struct DeviceScope {
  int prev = -1;
  explicit DeviceScope(int d) {
    cudaGetDevice(&prev);
    cudaSetDevice(d);
  }
  ~DeviceScope() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// sums of the 16 terms over the columns of one device: one block per term, a tree in shared memory.  (The order of the summation
// differs from a serial host loop; the totals are a diagnostic, not a result of the path.)
__global__ void group_totals_kernel(const double* __restrict__ sum, int64_t ncol, double* __restrict__ tot) {
  __shared__ double sh[256];
  const int k = blockIdx.x;
  double a = 0.0;
  for (int64_t c = threadIdx.x; c < ncol; c += blockDim.x) a += sum[c * 16 + k];
  sh[threadIdx.x] = a;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) tot[k] = sh[0];
}

void release_buffers(lgarb_group* g) {
  for (size_t k = 0; k < g->dev.size(); k++) {
    DeviceScope ds(g->dev[k]);
    if (k < g->d_sum.size() && g->d_sum[k]) cudaFree(g->d_sum[k]);
    if (k < g->d_tot.size() && g->d_tot[k]) cudaFree(g->d_tot[k]);
    if (k == 0) {
      if (g->d_all) cudaFree(g->d_all);
      if (g->d_tot_all) cudaFree(g->d_tot_all);
    }
  }
  g->d_sum.clear();
  g->d_tot.clear();
  g->d_all = g->d_tot_all = nullptr;
  for (ncclComm_t c : g->comm) ncclCommDestroy(c);
  g->comm.clear();
}

template <typename F>
int guarded(lgarb_group* g, F&& f) {
  if (!g) return LGARB_FAILURE;
  try {
    f();
    return LGARB_SUCCESS;
  } catch (const std::exception& ex) {
    g->err = ex.what();
    return LGARB_FAILURE;
  }
}

}  // namespace

extern "C" {

int lgarb_group_create(lgarb_group** out, int n_devices, const int* devices) {
  if (!out || n_devices < 1) return LGARB_FAILURE;
  int have = 0;
  if (cudaGetDeviceCount(&have) != cudaSuccess) have = 0;
  lgarb_group* g = new (std::nothrow) lgarb_group();
  if (!g) return LGARB_FAILURE;
  for (int k = 0; k < n_devices; k++) {
    const int d = devices ? devices[k] : k;
    if (d < 0 || d >= have || std::find(g->dev.begin(), g->dev.end(), d) != g->dev.end()) {
      delete g;
      return LGARB_FAILURE;
    }
    g->dev.push_back(d);
  }
  *out = g;
  return LGARB_SUCCESS;
}

int lgarb_group_finalize(lgarb_group* g) {
  return guarded(g, [&] {
    release_buffers(g);
    for (lgarb_engine* e : g->eng) {
      if (e) lgarb_destroy(e);
    }
    g->eng.clear();
    g->col0.clear();
    g->ncol.clear();
    g->initialized = false;
  });
}

int lgarb_group_destroy(lgarb_group* g) {
  if (!g) return LGARB_FAILURE;
  lgarb_group_finalize(g);
  delete g;
  return LGARB_SUCCESS;
}

const char* lgarb_group_last_error(const lgarb_group* g) { return g ? g->err.c_str() : "null group"; }

int lgarb_group_initialize(lgarb_group* g, const char* config_file, int64_t n_columns, const int32_t* layer_soil_type, const double* layer_thickness_cm) {
  return guarded(g, [&] {
    if (!config_file || n_columns < (int64_t)g->dev.size()) throw std::runtime_error("bad arguments (at least one column per device)");
    if ((layer_soil_type == nullptr) != (layer_thickness_cm == nullptr)) throw std::runtime_error("layer_soil_type and layer_thickness_cm come together");
    if (g->initialized) lgarb_group_finalize(g);
    const int n = (int)g->dev.size();
    const int64_t per = (n_columns + n - 1) / n;  // lgar_c_b200/sharding.py:shard_range
    g->ntot = n_columns;
    g->eng.assign(n, nullptr);
    g->col0.assign(n, 0);
    g->ncol.assign(n, 0);
    int L = 0;
    for (int k = 0; k < n; k++) {
      g->col0[k] = std::min<int64_t>(n_columns, (int64_t)k * per);
      g->ncol[k] = std::max<int64_t>(0, std::min<int64_t>(n_columns, g->col0[k] + per) - g->col0[k]);
      if (g->ncol[k] == 0) throw std::runtime_error("a device would own no column: use fewer devices");
      if (lgarb_create(&g->eng[k]) != LGARB_SUCCESS) throw std::runtime_error("lgarb_create failed");
      int rc;
      if (layer_soil_type) {
        if (k == 0) {  // the layer count comes from the configuration: ask a one-column engine on the first device
          lgarb_engine* probe = nullptr;
          if (lgarb_create(&probe) != LGARB_SUCCESS || lgarb_initialize(probe, config_file, 1, g->dev[0]) != LGARB_SUCCESS) {
            std::string m = probe ? lgarb_last_error(probe) : "lgarb_create failed";
            if (probe) lgarb_destroy(probe);
            throw std::runtime_error(m);
          }
          L = lgarb_get_num_layers(probe);
          lgarb_destroy(probe);
        }
        rc = lgarb_initialize_columns(g->eng[k], config_file, g->ncol[k], g->dev[k], layer_soil_type + g->col0[k] * L, layer_thickness_cm + g->col0[k] * L);
      } else {
        rc = lgarb_initialize(g->eng[k], config_file, g->ncol[k], g->dev[k]);
      }
      if (rc != LGARB_SUCCESS) throw std::runtime_error(std::string("device ") + std::to_string(g->dev[k]) + ": " + lgarb_last_error(g->eng[k]));
    }
    g->initialized = true;
  });
}

int lgarb_group_num_devices(const lgarb_group* g) { return g ? (int)g->dev.size() : -1; }
int64_t lgarb_group_num_columns(const lgarb_group* g) { return g ? g->ntot : -1; }

int lgarb_group_shard(const lgarb_group* g, int k, int64_t* col0, int64_t* ncol) {
  if (!g || !g->initialized || k < 0 || k >= (int)g->dev.size()) return LGARB_FAILURE;
  if (col0) *col0 = g->col0[k];
  if (ncol) *ncol = g->ncol[k];
  return LGARB_SUCCESS;
}

lgarb_engine* lgarb_group_engine(lgarb_group* g, int k) { return (g && g->initialized && k >= 0 && k < (int)g->eng.size()) ? g->eng[k] : nullptr; }

int lgarb_group_update_steps_host(lgarb_group* g, int64_t n_steps, const double* h_precip, const double* h_pet, int n_outputs,
                                  const char* const* output_names, double* const* h_outputs) {
  return guarded(g, [&] {
    if (!g->initialized) throw std::runtime_error("group not initialized");
    if (n_outputs < 0 || n_outputs > LGARB_NUM_OUTPUT_SCALARS) throw std::runtime_error("bad output count");
    const int n = (int)g->dev.size();
    std::vector<int> rc(n, LGARB_SUCCESS);
    std::vector<std::thread> th;
    for (int k = 0; k < n; k++) {
      th.emplace_back([&, k] {
        std::vector<double*> outs(n_outputs > 0 ? n_outputs : 1, nullptr);
        for (int i = 0; i < n_outputs; i++) outs[i] = h_outputs[i] + g->col0[k];
        rc[k] = lgarb_update_steps_host_pitched(g->eng[k], n_steps, h_precip + g->col0[k], h_pet + g->col0[k], g->ntot, n_outputs, output_names, outs.data(), g->ntot);
      });
    }
    for (std::thread& t : th) t.join();
    for (int k = 0; k < n; k++)
      if (rc[k] != LGARB_SUCCESS) throw std::runtime_error(std::string("device ") + std::to_string(g->dev[k]) + ": " + lgarb_last_error(g->eng[k]));
  });
}

int lgarb_group_gather_mass_balance(lgarb_group* g, double* h_out, double* totals16, double* gather_ms) {
  return guarded(g, [&] {
    if (!g->initialized) throw std::runtime_error("group not initialized");
    if (!totals16) throw std::runtime_error("null totals");
    const int n = (int)g->dev.size();
    if (g->d_sum.empty()) {  // buffers and communicators, once
      g->d_sum.assign(n, nullptr);
      g->d_tot.assign(n, nullptr);
      for (int k = 0; k < n; k++) {
        DeviceScope ds(g->dev[k]);
        GROUP_CUDA(cudaMalloc((void**)&g->d_sum[k], (size_t)g->ncol[k] * 16 * sizeof(double)));
        GROUP_CUDA(cudaMalloc((void**)&g->d_tot[k], 16 * sizeof(double)));
        if (k == 0) {
          GROUP_CUDA(cudaMalloc((void**)&g->d_all, (size_t)g->ntot * 16 * sizeof(double)));
          GROUP_CUDA(cudaMalloc((void**)&g->d_tot_all, 16 * sizeof(double)));
        }
      }
      if (n > 1) {
        g->comm.assign(n, nullptr);
        GROUP_NCCL(ncclCommInitAll(g->comm.data(), n, g->dev.data()));
      }
    }
    std::vector<cudaStream_t> st(n);
    for (int k = 0; k < n; k++) {  // per-column terms and their sums, behind each engine's own stream
      DeviceScope ds(g->dev[k]);
      st[k] = (cudaStream_t)lgarb_get_stream(g->eng[k]);
      if (lgarb_get_global_mass_balance_device(g->eng[k], g->d_sum[k]) != LGARB_SUCCESS) throw std::runtime_error(lgarb_last_error(g->eng[k]));
      group_totals_kernel<<<16, 256, 0, st[k]>>>(g->d_sum[k], g->ncol[k], g->d_tot[k]);
      GROUP_CUDA(cudaGetLastError());
    }
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    {
      DeviceScope ds(g->dev[0]);
      GROUP_CUDA(cudaEventCreate(&ev0));
      GROUP_CUDA(cudaEventCreate(&ev1));
      GROUP_CUDA(cudaEventRecord(ev0, st[0]));
      GROUP_CUDA(cudaMemcpyAsync(g->d_all, g->d_sum[0], (size_t)g->ncol[0] * 16 * sizeof(double), cudaMemcpyDeviceToDevice, st[0]));
    }
    if (n > 1) {
      GROUP_NCCL(ncclGroupStart());
      for (int k = 1; k < n; k++) {
        {
          DeviceScope ds(g->dev[k]);
          GROUP_NCCL(ncclSend(g->d_sum[k], (size_t)g->ncol[k] * 16, ncclDouble, 0, g->comm[k], st[k]));
        }
        DeviceScope ds(g->dev[0]);
        GROUP_NCCL(ncclRecv(g->d_all + (size_t)g->col0[k] * 16, (size_t)g->ncol[k] * 16, ncclDouble, k, g->comm[0], st[0]));
      }
      for (int k = 0; k < n; k++) {
        DeviceScope ds(g->dev[k]);
        GROUP_NCCL(ncclReduce(g->d_tot[k], k == 0 ? g->d_tot_all : g->d_tot[k], 16, ncclDouble, ncclSum, 0, g->comm[k], st[k]));
      }
      GROUP_NCCL(ncclGroupEnd());
    } else {
      DeviceScope ds(g->dev[0]);
      GROUP_CUDA(cudaMemcpyAsync(g->d_tot_all, g->d_tot[0], 16 * sizeof(double), cudaMemcpyDeviceToDevice, st[0]));
    }
    {
      DeviceScope ds(g->dev[0]);
      GROUP_CUDA(cudaEventRecord(ev1, st[0]));
      if (h_out) GROUP_CUDA(cudaMemcpyAsync(h_out, g->d_all, (size_t)g->ntot * 16 * sizeof(double), cudaMemcpyDeviceToHost, st[0]));
      GROUP_CUDA(cudaMemcpyAsync(totals16, g->d_tot_all, 16 * sizeof(double), cudaMemcpyDeviceToHost, st[0]));
    }
    for (int k = 0; k < n; k++) {
      DeviceScope ds(g->dev[k]);
      GROUP_CUDA(cudaStreamSynchronize(st[k]));
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev0, ev1);
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    if (gather_ms) *gather_ms = ms;
  });
}

}  // extern "C"
