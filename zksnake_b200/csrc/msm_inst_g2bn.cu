// msm_inst_g2bn.cu -- MSM / point kernels instantiated for one (curve, group); separate TU so the four compile in parallel.
#include "msm_host.cuh"
namespace zkb {
ZKB_MSM_INSTANTIATE(g2bn, fq2_bn, 254, ZKB_BN254, 2)
}
