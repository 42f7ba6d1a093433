// ntt_warp_inst_bls2.cu -- instantiation of the register-resident NTT passes (ntt_warp.cuh) for fr_bls, 4 elements per lane.
#include <cuda_runtime.h>
#define ZKB_NTT_WARP_INSTANTIATE
#include "ntt_warp.cuh"
namespace zkb {
template int ntt_warp_launch_el<fr_bls, 2>(const fr_bls*, fr_bls*, const NttPass&, uint32_t, size_t, size_t, const PowTable<fr_bls>&,
                                            const PreTables<fr_bls>&, const PowTable<fr_bls>&, const fr_bls&, void*);
}
