// msm_inst_g2bls.cu -- MSM / point kernels instantiated for one (curve, group); separate TU so the four compile in parallel.
#include "msm_host.cuh"
namespace zkb {
ZKB_MSM_INSTANTIATE(g2bls, fq2_bls, 255, ZKB_BLS12_381, 2)
}
