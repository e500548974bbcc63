/* include/lgarb_group.h -- several GPUs of one node behind one handle: liblgarb_group.so (host C++ over liblgarb.so + NCCL).
 *
 * The reference has no multi-device notion (one BmiLGAR object per column, include/bmi_lgar.hxx:30-184; a driver such as ngen
 * owns as many objects as it has catchments).  Columns share nothing, so the path shards by CONTIGUOUS COLUMN BLOCKS: device k of
 * n owns the columns [k * ceil(N / n), min(N, (k + 1) * ceil(N / n))) -- the same split as lgar_c_b200/sharding.py:shard_range,
 * which the one-process-per-GPU launcher (bench.py under torchrun) uses.  There is NO collective on the step path: every device
 * runs its own lgarb_engine on its block, driven by a host thread of its own.  What is exchanged is the end-of-run summary: the 16
 * terms of lgar_global_mass_balance (reference src/lgar.cxx:1162-1216) of every column are gathered onto the first device with
 * NCCL point-to-point sends and their totals are summed with ncclReduce, timed on the device.
 *
 * A library of its own (it links libnccl): processes that only load liblgarb.so never see NCCL.
 */
#ifndef LGARB_GROUP_H
#define LGARB_GROUP_H

#include "lgarb.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lgarb_group lgarb_group;

/* devices: n_devices CUDA device indices (NULL: 0 .. n_devices - 1).  Fails if a device does not exist. */
int lgarb_group_create(lgarb_group** out, int n_devices, const int* devices);
int lgarb_group_destroy(lgarb_group* g);
const char* lgarb_group_last_error(const lgarb_group* g);

/* lgarb_initialize_columns on every device for its block of the n_columns columns (layer_soil_type / layer_thickness_cm:
 * [n_columns][num_layers] or NULL for the configuration's own, as there). */
int lgarb_group_initialize(lgarb_group* g, const char* config_file, int64_t n_columns, const int32_t* layer_soil_type,
                           const double* layer_thickness_cm);
int lgarb_group_finalize(lgarb_group* g);

int lgarb_group_num_devices(const lgarb_group* g);
int64_t lgarb_group_num_columns(const lgarb_group* g);
/* the block of device k: first column and number of columns; the engine that owns it (every verb of lgarb.h works on it) */
int lgarb_group_shard(const lgarb_group* g, int k, int64_t* col0, int64_t* ncol);
lgarb_engine* lgarb_group_engine(lgarb_group* g, int k);

/* lgarb_update_steps_host over GLOBAL host arrays [n_steps][n_columns] (forcing in, chosen per-step outputs back): one host
 * thread per device, each working on its block of columns (lgarb_update_steps_host_pitched).  Returns when every device is done. */
int lgarb_group_update_steps_host(lgarb_group* g, int64_t n_steps, const double* h_precip_mm_per_h, const double* h_pet_mm_per_h,
                                  int n_outputs, const char* const* output_names, double* const* h_outputs);

/* End-of-run gather.  Every device computes the 16 balance terms of its columns (lgarb_get_global_mass_balance_device) and their
 * sums; the per-column terms travel to the first device with ncclSend / ncclRecv, the sums with ncclReduce; then one copy to the
 * host.  h_out: [n_columns][16] in column order, or NULL (totals only); totals16: sums over all columns; gather_ms (nullable):
 * device time of the NCCL part on the first device. */
int lgarb_group_gather_mass_balance(lgarb_group* g, double* h_out, double* totals16, double* gather_ms);

#ifdef __cplusplus
}
#endif
#endif
