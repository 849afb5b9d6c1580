// Minimal translation unit for tools/sass/mix.sh: k_var_base as in pa_kernels.cuh (6 resident blocks), plain loaders.
#include <stddef.h>
#include "pa_smul.cuh"
// minimal copies of the loaders for a SASS experiment
__device__ __forceinline__ void ldp(jac &P, const unsigned char *p){ fe_from_be(P.X,p); fe_from_be(P.Y,p+32); fe_set_one(P.Z);} 
__global__ void __launch_bounds__(128, 6)
k_var_base(const unsigned char *points, const unsigned char *scalars, u32 *jout, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  jac P, r; sc k;
  ldp(P, points + 64 * (size_t)i);
  for (int j=0;j<8;++j) k.v[j] = ((const u32*)scalars)[8*(size_t)i+j];
  var_base_mul(r, P, k);
  for (int j=0;j<8;++j){ jout[24*(size_t)i+j]=r.X.v[j]; jout[24*(size_t)i+8+j]=r.Y.v[j]; jout[24*(size_t)i+16+j]=r.Z.v[j];}
}
