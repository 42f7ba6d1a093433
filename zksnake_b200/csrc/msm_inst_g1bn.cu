// msm_inst_g1bn.cu -- MSM / point kernels instantiated for one (curve, group); separate TU so the four compile in parallel.
#include "msm_host.cuh"
namespace zkb {
ZKB_MSM_INSTANTIATE(g1bn, fq_bn, 254, ZKB_BN254, 1)
}
