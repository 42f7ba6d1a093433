// msm_common.cu -- (curve, group) dispatch for the MSM and point-vector entry points.
#include <cuda_runtime.h>
#include "zkb_internal.h"

namespace zkb {

int g_msm_c = 0, g_msm_seg = 0, g_msm_kchunk = 0;
void msm_set_tuning(int c, int seg, int kchunk) {
  g_msm_c = c;
  g_msm_seg = seg;
  g_msm_kchunk = kchunk;
}

#define DECL(SUFFIX)                                                                      \
  int msm_run_##SUFFIX(const void* p, const void* s, size_t n, uint64_t* o, int* inf);  \
  int points_conv_##SUFFIX(int to, size_t n, void* p);                                   \
  int batch_mul_##SUFFIX(const void* b, int single, const void* s, size_t n, void* o);
DECL(g1bn) DECL(g2bn) DECL(g1bls) DECL(g2bls)

#define DISPATCH(CALL_BN1, CALL_BN2, CALL_BL1, CALL_BL2)                  \
  if (curve == ZKB_BN254 && group == 1) return CALL_BN1;                  \
  if (curve == ZKB_BN254 && group == 2) return CALL_BN2;                  \
  if (curve == ZKB_BLS12_381 && group == 1) return CALL_BL1;              \
  if (curve == ZKB_BLS12_381 && group == 2) return CALL_BL2;              \
  return set_error(ZKB_ERR_ARG, "unknown (curve, group)");

int msm_dev(int curve, int group, const void* p, const void* s, size_t n, uint64_t* o, int* inf) {
  DISPATCH(msm_run_g1bn(p, s, n, o, inf), msm_run_g2bn(p, s, n, o, inf), msm_run_g1bls(p, s, n, o, inf),
           msm_run_g2bls(p, s, n, o, inf))
}
int points_to_mont_dev(int curve, int group, size_t n, void* p) {
  DISPATCH(points_conv_g1bn(1, n, p), points_conv_g2bn(1, n, p), points_conv_g1bls(1, n, p), points_conv_g2bls(1, n, p))
}
int points_from_mont_dev(int curve, int group, size_t n, void* p) {
  DISPATCH(points_conv_g1bn(0, n, p), points_conv_g2bn(0, n, p), points_conv_g1bls(0, n, p), points_conv_g2bls(0, n, p))
}
int batch_mul_dev(int curve, int group, const void* b, int single, const void* s, size_t n, void* o) {
  DISPATCH(batch_mul_g1bn(b, single, s, n, o), batch_mul_g2bn(b, single, s, n, o), batch_mul_g1bls(b, single, s, n, o),
           batch_mul_g2bls(b, single, s, n, o))
}

}  // namespace zkb
