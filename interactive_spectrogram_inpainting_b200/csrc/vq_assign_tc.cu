// tcgen05 (3xTF32) nearest-code search -- placeholder until the kernel lands.
#include "common.cuh"

namespace isi {

bool assign_tc_supported(const isi_rows_layout&, int64_t, int, int) { return false; }

int launch_prepare_tc(const float*, int, int, const Prepared&, cudaStream_t) { return ISI_OK; }

int launch_assign_tc(const float*, const isi_rows_layout&, int64_t, int, int, const Prepared&,
                     int64_t*, float*, cudaStream_t) {
  return ISI_ERR_UNSUPPORTED;
}

}  // namespace isi
