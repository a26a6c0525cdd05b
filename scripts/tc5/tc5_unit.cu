// Unit test of the tcgen05 building block used by the overlap kernel's MODE 6:
//   D[128 x 64] = A[128 x 16] * B[16 x 64]  as 3xTF32 (lo.hi + hi.lo + hi.hi) with
//   tcgen05.mma.cta_group::1.kind::tf32, operands in shared memory in the no-swizzle
//   K-major canonical layout (core matrix = 8 rows x 16 bytes), accumulator in TMEM,
//   read back with tcgen05.ld.32x32b.  Checked against fp64 on the host.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc5_unit tc5_unit.cu
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
// byte offset of element (row, k) of a K-major operand tile: 8-row groups SBO apart, 16-byte k chunks LBO apart
__device__ __forceinline__ uint32_t kmaj_off(int row, int k, uint32_t sbo, uint32_t lbo) {
  return (uint32_t)(row >> 3) * sbo + (uint32_t)(k >> 2) * lbo + (uint32_t)(row & 7) * 16u + (uint32_t)(k & 3) * 4u;
}
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fffu);
  d |= (uint64_t)((lbo >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (sm_100)
  return d;                // layout type 0: no swizzle
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}

constexpr int M = 128, N = 64, K = 16;
constexpr uint32_t LBO = 128, SBO_A = 512, SBO_B = 512;

__global__ void __launch_bounds__(128, 1) tc5_unit_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                         float* __restrict__ D, int* __restrict__ status) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sAhi = smem;                    // 128 x 16 fp32 = 8 KB
  unsigned char* sAlo = smem + 8192;
  unsigned char* sBhi = smem + 16384;            // 64 x 16 fp32 = 4 KB   (B stored as [n][k])
  unsigned char* sBlo = smem + 20480;
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int e = tid; e < M * K; e += 128) {
    const int r = e / K, k = e % K;
    uint32_t hi, lo;
    split_tf32(A[e], hi, lo);
    *reinterpret_cast<uint32_t*>(sAhi + kmaj_off(r, k, SBO_A, LBO)) = hi;
    *reinterpret_cast<uint32_t*>(sAlo + kmaj_off(r, k, SBO_A, LBO)) = lo;
  }
  for (int e = tid; e < K * N; e += 128) {
    const int k = e / N, n = e % N;  // B is [k][n] row-major in global memory
    uint32_t hi, lo;
    split_tf32(B[e], hi, lo);
    *reinterpret_cast<uint32_t*>(sBhi + kmaj_off(n, k, SBO_B, LBO)) = hi;
    *reinterpret_cast<uint32_t*>(sBlo + kmaj_off(n, k, SBO_B, LBO)) = lo;
  }
  if (tid == 0) {
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&mbar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(
        (unsigned)__cvta_generic_to_shared(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // generic-proxy writes of the operands -> visible to the async proxy (the tensor core reads smem through it)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;

  if (tid == 0) {
    // instruction descriptor: D = f32 (bit 4), A = B = TF32 (2 << 7, 2 << 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t aH = (uint32_t)__cvta_generic_to_shared(sAhi), aL = (uint32_t)__cvta_generic_to_shared(sAlo);
    const uint32_t bH = (uint32_t)__cvta_generic_to_shared(sBhi), bL = (uint32_t)__cvta_generic_to_shared(sBlo);
    uint32_t acc = 0;
#pragma unroll
    for (int ks = 0; ks < K / 8; ++ks) {
      const uint32_t o = (uint32_t)ks * 2u * LBO;  // 8 tf32 = two 16-byte chunks further along K
      umma_tf32(tmem, smem_desc(aL + o, LBO, SBO_A), smem_desc(bH + o, LBO, SBO_B), idesc, acc); acc = 1;
      umma_tf32(tmem, smem_desc(aH + o, LBO, SBO_A), smem_desc(bL + o, LBO, SBO_B), idesc, acc);
      umma_tf32(tmem, smem_desc(aH + o, LBO, SBO_A), smem_desc(bH + o, LBO, SBO_B), idesc, acc);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
        (unsigned)__cvta_generic_to_shared(&mbar)) : "memory");
  }
  {
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&mbar);
    unsigned done = 0;
    for (int spin = 0; !done; ++spin) {
      if (spin > (1 << 22)) { if (lane == 0) atomicExch(status, 2); break; }
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(bar) : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // warp w reads TMEM lanes 32 w .. 32 w + 31 (= rows of D), 64 columns
  uint32_t v[64];
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll
  for (int c0 = 0; c0 < 64; c0 += 16) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[c0 + 0]), "=r"(v[c0 + 1]), "=r"(v[c0 + 2]), "=r"(v[c0 + 3]), "=r"(v[c0 + 4]), "=r"(v[c0 + 5]),
          "=r"(v[c0 + 6]), "=r"(v[c0 + 7]), "=r"(v[c0 + 8]), "=r"(v[c0 + 9]), "=r"(v[c0 + 10]), "=r"(v[c0 + 11]),
          "=r"(v[c0 + 12]), "=r"(v[c0 + 13]), "=r"(v[c0 + 14]), "=r"(v[c0 + 15])
        : "r"(taddr + (uint32_t)c0));
  }
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  const int row = warp * 32 + lane;
#pragma unroll
  for (int c = 0; c < 64; ++c) D[row * N + c] = __uint_as_float(v[c]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem));
  if (tid == 0) atomicCAS(status, 0, 1);
}

int main() {
  std::vector<float> hA(M * K), hB(K * N), hD(M * N);
  srand(1);
  for (auto& x : hA) x = (float)rand() / RAND_MAX - 0.5f;
  for (auto& x : hB) x = (float)rand() / RAND_MAX - 0.5f;
  float *dA, *dB, *dD; int* dS;
  CK(cudaMalloc(&dA, hA.size() * 4)); CK(cudaMalloc(&dB, hB.size() * 4)); CK(cudaMalloc(&dD, hD.size() * 4));
  CK(cudaMalloc(&dS, 4)); CK(cudaMemset(dS, 0, 4)); CK(cudaMemset(dD, 0, hD.size() * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(tc5_unit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
  tc5_unit_kernel<<<1, 128, 24576 + 1024>>>(dA, dB, dD, dS);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  int st = 0;
  CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  for (int r = 0; r < M; ++r)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)hA[r * K + k] * (double)hB[k * N + n];
      maxerr = fmax(maxerr, fabs(ref - hD[r * N + n]));
      maxref = fmax(maxref, fabs(ref));
    }
  printf("tc5_unit: status %d (1 = ok, 2 = mbarrier timeout), max abs err %.3e, max |ref| %.3f, rel %.3e\n", st, maxerr, maxref,
         maxerr / maxref);
  return (st == 1 && maxerr / maxref < 1e-5) ? 0 : 1;
}
