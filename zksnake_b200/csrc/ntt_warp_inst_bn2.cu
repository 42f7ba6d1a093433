// ntt_warp_inst_bn2.cu -- instantiation of the register-resident NTT passes (ntt_warp.cuh) for fr_bn, 4 elements per lane.
#include <cuda_runtime.h>
#define ZKB_NTT_WARP_INSTANTIATE
#include "ntt_warp.cuh"
namespace zkb {
template int ntt_warp_launch_el<fr_bn, 2>(const fr_bn*, fr_bn*, const NttPass&, uint32_t, size_t, size_t, const PowTable<fr_bn>&,
                                            const PreTables<fr_bn>&, const PowTable<fr_bn>&, const fr_bn&, void*);
}
