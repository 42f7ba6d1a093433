// msm_inst_g1bn.cu -- MSM / point kernels instantiated for one (curve, group); separate TU so the four compile in parallel.
#include "msm_host.cuh"
namespace zkb {
ZKB_MSM_INSTANTIATE(g1bn, fq_bn, 254, ZKB_BN254, 1)
}

namespace zkb {
// the plan heuristics live in msm_host.cuh; this translation unit exports them for zkb_groth16_pk_msm_info
void msm_plan_info(size_t n, uint32_t scalar_bits, uint32_t wworld, uint32_t* c, uint32_t* W) {
  MsmPlan pl = msm_make_plan(n, scalar_bits, 0, wworld ? wworld : 1);
  *c = pl.c;
  *W = pl.nwin_total;
}
template <class F>
static void kernel_info_t(int* lanes, int* ctas) {
  *lanes = AccumField<F>::PAIR ? 2 : 1;
  *ctas = AccumField<F>::MINB;
}
void msm_kernel_info(int curve, int group, int* lanes_per_point, int* ctas_per_sm) {
  if (curve == ZKB_BN254) {
    if (group == 2) kernel_info_t<fq2_bn>(lanes_per_point, ctas_per_sm);
    else kernel_info_t<fq_bn>(lanes_per_point, ctas_per_sm);
  } else {
    if (group == 2) kernel_info_t<fq2_bls>(lanes_per_point, ctas_per_sm);
    else kernel_info_t<fq_bls>(lanes_per_point, ctas_per_sm);
  }
}
}  // namespace zkb
