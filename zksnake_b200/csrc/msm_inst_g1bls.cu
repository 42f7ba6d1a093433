// msm_inst_g1bls.cu -- MSM / point kernels instantiated for one (curve, group); separate TU so the four compile in parallel.
#include "msm_host.cuh"
namespace zkb {
ZKB_MSM_INSTANTIATE(g1bls, fq_bls, 255, ZKB_BLS12_381, 1)
}
