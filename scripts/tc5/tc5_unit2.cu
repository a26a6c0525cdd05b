// Unit test 2: the MODE 6 flow of the overlap kernel in isolation.
//   per CTA (8 warps = 8 images): Y_i (32 x 32, per image) -> TMEM as the A operand (hi / lo, tcgen05.st, lane = row),
//   B = W (32 x 32, shared by all images) in smem (K-major, no swizzle, hi / lo),
//   D_i = Y_i W as 3xTF32 tcgen05.mma (A from TMEM), two M = 128 tiles (images 0-3, 4-7),
//   D back with tcgen05.ld (lane = row).  Repeated ITER times with a CTA barrier + mbarrier per round to
//   measure the exposed latency of one round (clock64).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi)) + 0x1000u;
}
__device__ __forceinline__ uint32_t kmaj_off(int row, int k, uint32_t sbo, uint32_t lbo) {
  return (uint32_t)(row >> 3) * sbo + (uint32_t)(k >> 2) * lbo + (uint32_t)(row & 7) * 16u + (uint32_t)(k & 3) * 4u;
}
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
#define ST32(addr, v)                                                                                                   \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"  \
               "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(addr),                          \
               "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),        \
               "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]),            \
               "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]),           \
               "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])                        \
               : "memory")
#define LD32(addr, v)                                                                                                   \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"     \
               "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                                    \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),         \
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),   \
                 "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),              \
                 "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),              \
                 "=r"(v[30]), "=r"(v[31])                                                                                \
               : "r"(addr))

constexpr int KD = 32, ND = 32;
constexpr uint32_t LBO = 128, SBO = 1024;
// TMEM columns: A hi tile t at 64 t, A lo at 64 t + 32 (t = 0, 1); D tile t at 128 + 32 t
constexpr int ITER = 200;

__global__ void __launch_bounds__(256, 1) tc5_unit2_kernel(const float* __restrict__ Y, const float* __restrict__ W,
                                                          float* __restrict__ D, int* __restrict__ status,
                                                          long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sBhi = smem;         // [n][k] 32 x 32 fp32 = 4 KB
  unsigned char* sBlo = smem + 4096;
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int e = tid; e < KD * ND; e += 256) {
    const int k = e / ND, n = e % ND;  // W is [k][n]
    uint32_t hi, lo;
    split_tf32(W[e], hi, lo);
    *reinterpret_cast<uint32_t*>(sBhi + kmaj_off(n, k, SBO, LBO)) = hi;
    *reinterpret_cast<uint32_t*>(sBlo + kmaj_off(n, k, SBO, LBO)) = lo;
  }
  if (tid == 0) {
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&mbar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(
        (unsigned)__cvta_generic_to_shared(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  const int tile = warp >> 2, q = warp & 3;          // image = warp: M tile `tile`, lane quarter q
  const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
  // this lane's row of Y_image (row = lane), 32 values along K
  uint32_t yh[32], yl[32], d[32];
  {
    const float* yr = Y + ((size_t)warp * 32 + lane) * KD;
#pragma unroll
    for (int k = 0; k < 32; ++k) split_tf32(yr[k], yh[k], yl[k]);
  }
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(ND >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t bH = (uint32_t)__cvta_generic_to_shared(sBhi), bL = (uint32_t)__cvta_generic_to_shared(sBlo);
  const unsigned bar = (unsigned)__cvta_generic_to_shared(&mbar);
  bool bad = false;
  const long long t0 = clock64();
  for (int it = 0; it < ITER && !bad; ++it) {
    ST32(lane_addr + (uint32_t)(64 * tile), yh);
    ST32(lane_addr + (uint32_t)(64 * tile + 32), yl);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const uint32_t dcol = tmem + 128u + 32u * t, ah = tmem + 64u * t, al = ah + 32u;
#pragma unroll
        for (int ks = 0; ks < KD / 8; ++ks) {
          const uint32_t o = (uint32_t)ks * 2u * LBO;
          umma_tf32_ts(dcol, al + 8u * ks, smem_desc(bH + o, LBO, SBO), idesc, ks > 0);
          umma_tf32_ts(dcol, ah + 8u * ks, smem_desc(bL + o, LBO, SBO), idesc, 1);
          umma_tf32_ts(dcol, ah + 8u * ks, smem_desc(bH + o, LBO, SBO), idesc, 1);
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    }
    {
      unsigned done = 0;
      const unsigned parity = (unsigned)(it & 1);
      for (int spin = 0; !done; ++spin) {
        if (spin > (1 << 22)) { if (lane == 0) atomicExch(status, 2); bad = true; break; }
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
      }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    LD32(lane_addr + 128u + 32u * tile, d);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  }
  const long long t1 = clock64();
  {
    float* dr = D + ((size_t)warp * 32 + lane) * ND;
#pragma unroll
    for (int n = 0; n < 32; ++n) dr[n] = __uint_as_float(d[n]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
  if (tid == 0) { atomicCAS(status, 0, 1); cycles[0] = (t1 - t0) / ITER; }
}

int main() {
  std::vector<float> hY(8 * 32 * KD), hW(KD * ND), hD(8 * 32 * ND);
  srand(2);
  for (auto& x : hY) x = (float)rand() / RAND_MAX - 0.5f;
  for (auto& x : hW) x = (float)rand() / RAND_MAX - 0.5f;
  float *dY, *dW, *dD; int* dS; long long* dC;
  CK(cudaMalloc(&dY, hY.size() * 4)); CK(cudaMalloc(&dW, hW.size() * 4)); CK(cudaMalloc(&dD, hD.size() * 4));
  CK(cudaMalloc(&dS, 4)); CK(cudaMemset(dS, 0, 4)); CK(cudaMalloc(&dC, 8)); CK(cudaMemset(dD, 0, hD.size() * 4));
  CK(cudaMemcpy(dY, hY.data(), hY.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dW, hW.data(), hW.size() * 4, cudaMemcpyHostToDevice));
  tc5_unit2_kernel<<<1, 256, 8192 + 1024>>>(dY, dW, dD, dS, dC);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  int st = 0; long long cyc = 0;
  CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  for (int r = 0; r < 256; ++r)
    for (int n = 0; n < ND; ++n) {
      double ref = 0;
      for (int k = 0; k < KD; ++k) ref += (double)hY[r * KD + k] * (double)hW[k * ND + n];
      maxerr = fmax(maxerr, fabs(ref - hD[r * ND + n]));
      maxref = fmax(maxref, fabs(ref));
    }
  printf("tc5_unit2: status %d (1 = ok), rel err %.3e, cycles per round (st + barrier + 24 MMA + commit + wait + ld) %lld\n", st,
         maxerr / maxref, cyc);
  return (st == 1 && maxerr / maxref < 1e-5) ? 0 : 1;
}
