// ufair_sampler.cu -- on-device ensemble sampler (include/ufair.h, "on-device ensemble sampler"):
// Philox4x32-10 counter-based stream keyed by (seed, global member index), Box-Muller normals,
// lognormal / normal perturbations of a base parameter table, pool fractions renormalised.
// One thread per (member, gas) -- 9 Philox blocks = 18 normals = the gas's 17 rows + its emission
// scale -- plus one thread per member for the thermal rows and the scenario index.  Every store is
// coalesced over the member axis.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ufair.h"
#include "ufair_internal.h"

namespace ufair {

struct U4 {
  uint32_t x, y, z, w;
};

__device__ __forceinline__ U4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return U4{c0, c1, c2, c3};
}

constexpr uint32_t kStreamTag = 0x55464152u;  // "UFAR"

__device__ __forceinline__ double unit_open(uint32_t lo, uint32_t hi) {  // (0, 1), 53 bits
  const uint64_t v = ((uint64_t)hi << 32) | lo;
  return ((double)(v >> 11) + 0.5) * 0x1.0p-53;
}

// the two normals of one Philox block
__device__ __forceinline__ void normal_pair(uint64_t seed, uint64_t member, uint32_t block, double* z0, double* z1) {
  const U4 x = philox4x32_10((uint32_t)member, (uint32_t)(member >> 32), block, kStreamTag, (uint32_t)seed,
                             (uint32_t)(seed >> 32));
  const double ua = unit_open(x.x, x.y), ub = unit_open(x.z, x.w);
  const double r = sqrt(-2.0 * log(ua));
  double s, c;
  sincospi(2.0 * ub, &s, &c);
  *z0 = r * c;
  *z1 = r * s;
}

__device__ __forceinline__ double perturb(double base, double sigma, int dist, double z) {
  if (dist == UFAIR_DIST_LOGNORMAL) return base * exp(sigma * z);
  if (dist == UFAIR_DIST_NORMAL) return base * (1.0 + sigma * z);
  return base;
}

template <typename Real>
__global__ void __launch_bounds__(256)
    sample_kernel(const __grid_constant__ ufair_sampler sp, long long first, long long n, long long ld, Real* gp, Real* tp,
                  Real* esc, int32_t* scen) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t member = (uint64_t)(first + i);
  const int g = blockIdx.y;
  if (g < sp.n_gas) {
    double z[18];
#pragma unroll
    for (int b = 0; b < 9; ++b) normal_pair(sp.seed, member, (uint32_t)(9 * g + b), &z[2 * b], &z[2 * b + 1]);
    double v[UFAIR_GP_COUNT];
#pragma unroll
    for (int r = 0; r < UFAIR_GP_COUNT; ++r) v[r] = perturb(sp.gas_base[g][r], sp.gas_sigma[g][r], sp.gas_dist[g][r], z[r]);
    const double sa = ((v[UFAIR_GP_A0] + v[UFAIR_GP_A0 + 1]) + v[UFAIR_GP_A0 + 2]) + v[UFAIR_GP_A0 + 3];
    if (sa > 0.0) {
#pragma unroll
      for (int q = 0; q < 4; ++q) v[UFAIR_GP_A0 + q] = v[UFAIR_GP_A0 + q] / sa;
    }
    if (gp) {
#pragma unroll
      for (int r = 0; r < UFAIR_GP_COUNT; ++r) gp[((long long)g * UFAIR_GP_COUNT + r) * ld + i] = (Real)v[r];
    }
    if (esc) esc[(long long)g * ld + i] = (Real)(1.0 + sp.e_scale_sigma * z[17]);
  } else {
    if (tp) {
      double z[4];
      normal_pair(sp.seed, member, 36u, &z[0], &z[1]);
      normal_pair(sp.seed, member, 37u, &z[2], &z[3]);
#pragma unroll
      for (int k = 0; k < UFAIR_TP_COUNT; ++k)
        tp[(long long)k * ld + i] = (Real)perturb(sp.thermal_base[k], sp.thermal_sigma[k], sp.thermal_dist[k], z[k]);
    }
    if (scen) {
      const U4 x = philox4x32_10((uint32_t)member, (uint32_t)(member >> 32), 38u, kStreamTag, (uint32_t)sp.seed,
                                 (uint32_t)(sp.seed >> 32));
      scen[i] = (int32_t)(((uint64_t)x.x * (uint64_t)(uint32_t)sp.n_scen) >> 32);
    }
  }
}

template <typename Real>
static int sample(const ufair_sampler* s, int64_t first, int64_t n, int64_t ld, Real* gp, Real* tp, Real* esc,
                  int32_t* scen, cudaStream_t stream) {
  if (!s) return set_error(UFAIR_ERR_ARG, "sampler is NULL");
  if (s->struct_size != sizeof(ufair_sampler))
    return set_error(UFAIR_ERR_ARG, "struct_size %u != sizeof(ufair_sampler) %zu (header/binding mismatch)", s->struct_size,
                     sizeof(ufair_sampler));
  if (s->n_gas < 1 || s->n_gas > UFAIR_MAX_GAS) return set_error(UFAIR_ERR_ARG, "n_gas %d outside 1..%d", s->n_gas, UFAIR_MAX_GAS);
  if (s->n_scen < 1) return set_error(UFAIR_ERR_ARG, "n_scen must be >= 1");
  if (first < 0 || n < 0 || ld < n) return set_error(UFAIR_ERR_ARG, "bad first_member / n_member / ld_member");
  for (int g = 0; g < s->n_gas; ++g)
    for (int r = 0; r < UFAIR_GP_COUNT; ++r)
      if (s->gas_dist[g][r] > UFAIR_DIST_NORMAL) return set_error(UFAIR_ERR_ARG, "gas_dist[%d][%d] is not a UFAIR_DIST_*", g, r);
  for (int k = 0; k < UFAIR_TP_COUNT; ++k)
    if (s->thermal_dist[k] > UFAIR_DIST_NORMAL) return set_error(UFAIR_ERR_ARG, "thermal_dist[%d] is not a UFAIR_DIST_*", k);
  if (n == 0) return UFAIR_OK;
  const dim3 grid((unsigned)((n + 255) / 256), (unsigned)(s->n_gas + 1));
  sample_kernel<Real><<<grid, 256, 0, stream>>>(*s, first, n, ld, gp, tp, esc, scen);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? UFAIR_OK : cuda_error(e, "sample_kernel launch");
}

}  // namespace ufair

extern "C" {
int ufair_sample_f64(const ufair_sampler* s, int64_t first_member, int64_t n_member, int64_t ld_member, double* gas_params,
                     double* thermal_params, double* e_scale, int32_t* scen_idx, void* stream) {
  return ufair::sample<double>(s, first_member, n_member, ld_member, gas_params, thermal_params, e_scale, scen_idx,
                               (cudaStream_t)stream);
}
int ufair_sample_f32(const ufair_sampler* s, int64_t first_member, int64_t n_member, int64_t ld_member, float* gas_params,
                     float* thermal_params, float* e_scale, int32_t* scen_idx, void* stream) {
  return ufair::sample<float>(s, first_member, n_member, ld_member, gas_params, thermal_params, e_scale, scen_idx,
                              (cudaStream_t)stream);
}
}
