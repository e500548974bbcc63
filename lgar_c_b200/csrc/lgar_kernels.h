// lgar_c_b200/csrc/lgar_kernels.h -- launch interface between the host engine and the step kernels.
#ifndef LGAR_KERNELS_H
#define LGAR_KERNELS_H

#ifndef LGAR_HOSTSIM
#include <cuda_runtime.h>
#endif

#include "lgar_types.h"

#define LGAR_ACTIVE_THREADS 128

// optional per-step output sinks: sinks.p[k] (device, may be null) receives output k of this step
// for every column, so a multi-step run can stream chosen variables into a [step][column] series.
struct LgarSinks {
  double* p[LGAR_NOUT];
};

#ifndef LGAR_HOSTSIM
// ev (nullable): three events recorded before the stream kernel, between the kernels, after the active kernel
cudaError_t lgar_launch_step(const LgarConfig& cfg, const LgarDeviceState& d, const LgarSinks& sinks, int active_blocks,
                             int force_all_active, cudaStream_t stream, cudaEvent_t* ev);
// sum of d.nfront over all columns, added into *acc (unsigned long long, device)
cudaError_t lgar_launch_sum_fronts(const LgarDeviceState& d, unsigned long long* acc, cudaStream_t stream);
cudaError_t lgar_launch_gather_fronts(const LgarDeviceState& d, int which, double scale, double* dst, int64_t col0, int64_t ncol,
                                      cudaStream_t stream);
cudaError_t lgar_launch_gather_layers(const LgarDeviceState& d, double* dst, int64_t col0, int64_t ncol, int L, cudaStream_t stream);
int lgar_active_kernel_max_blocks_per_sm();
#endif

#endif
