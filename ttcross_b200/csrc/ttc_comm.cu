// =============================================================================
// ttc_comm.cu — multi-GPU communicator of the TT-cross sweep (NCCL, loaded lazily).
// Placeholder until the block partition across GPUs lands: the entry points exist
// so that the C-ABI is complete and fail loudly.
// =============================================================================
#include "../../include/ttcross_b200.h"

extern "C" {
int ttc_comm_unique_id(void* id128) { (void)id128; return TTC_ERR_COMM; }
int ttc_comm_init(ttc_handle* h, int nranks, int rank, const void* id128) {
    (void)h; (void)nranks; (void)rank; (void)id128;
    return TTC_ERR_COMM;
}
}
