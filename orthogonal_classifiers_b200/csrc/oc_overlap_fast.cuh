// Specialised overlap kernel (fp32): per image, left environment up to the
// hairy site c, right environment down to it, centre contraction with the
// label leg open (experiments.py:332-337), |.| and argmax fused.
//
// Persistent CTAs (one per SM).  A tiny prep kernel re-lays the classifier out
// once per call:
//   * every non-hairy site in the order the sweeps consume it — left sites
//     [s][B][B'], right sites transposed [s][B'][B], rows padded to a multiple of
//     4 floats — as ONE contiguous block that each CTA pulls into shared memory
//     with a single bulk TMA copy (cp.async.bulk, SASS UBLKCP) + mbarrier;
//   * the hairy site as Wc[(s,B,B')][label] so that the 16 label sums read it
//     with coalesced 128-bit loads (it stays L1/L2 resident: 64 KB at D = 32).
// One warp per image; every contraction is C[m][n] = sum_k Xt[k][m] Y[k][n] with
// BOTH operands k-major in shared memory, so the inner loop is two LDS.128 and
// sixteen FFMA per lane (4x4 register tile).
#pragma once
#include "oc_generic.cuh"

namespace oc {

constexpr int kOvMaxSmallW = 12288;  // floats of non-hairy classifier sites kept in smem
constexpr int kOvMaxA = 3072;        // floats of one image's (padded) MPS in smem
constexpr int kOvE = 1024;           // 32 x 32 environment
constexpr int kOvT = 2048;           // 2 x 32 x 32 scratch
constexpr int kOvR = 256;            // 16 x 16 right environment (x2, ping-pong)
constexpr int kOvPre = 22;           // float4 prefetch registers per lane: images of <= 22*128 floats
// floats of per-warp scratch; the +4 puts the 8 warps' buffers on different banks (the CTA-wide
// label sums read element k of all 8 H buffers in one fragment load)
constexpr int kOvPerWarp = kOvMaxA + kOvE + kOvT + 2 * kOvR + 4;

__host__ __device__ inline int r4(int x) { return (x + 3) & ~3; }

struct OverlapFast {
  int64_t w_off[OC_MAX_SITES];   // offset of site n inside the small-W block (floats); -1 for the hairy site
  int64_t a_off[OC_MAX_SITES];   // offset of site n inside the per-warp MPS staging
  int small_w_total;             // floats in the small-W block (multiple of 4)
  int a_total;                   // floats of the per-warp MPS staging
  int Spad;                      // label leg padded to a power of two >= 4
  int64_t wc_elems;              // floats of the re-laid hairy site: 2 * B_c * r4(B_{c+1}) * Spad
  int w4u_off;                   // MODE 6: offset (floats) of site 4 as a tcgen05 B operand inside the small-W
                                 // block (hi tile, then lo tile, 1024 floats each); -1: absent
};

// ---- MODE 6 (tcgen05): site 4 of the natural 32/32 profile as ONE shared-operand product per CTA ----
// E5[(image, a')][B'] = sum_(s,B) Y[(image, a')][(s,B)] W4[(s,B)][B'],  Y = A4^T E4 per image (mma.sync):
// the classifier site is the B operand of a tcgen05.mma (M = 128 = 4 images x 32 rows, N = 32, K = 32,
// kind::tf32, 3xTF32), resident in shared memory in the no-swizzle K-major canonical layout
// (core matrix = 8 rows x 16 bytes; LBO between 16-byte k chunks, SBO between 8-row groups);
// the A operand (Y, hi / lo) and the accumulator live in tensor memory.
constexpr int kTc5MaxA = 2752;                     // per-warp MPS staging of the natural plan (2,736 floats)
constexpr int kTc5PerWarp = kTc5MaxA + 1024 + 2048 + 2 * 256 + 4;
constexpr uint32_t kTc5Lbo = 128, kTc5Sbo = 1024;  // bytes
constexpr int kTc5TileFloats = 1024;               // 32 x 32 fp32
constexpr uint32_t kTc5Cols = 256;                 // TMEM columns: A hi/lo of tile t at 64 t / 64 t + 32, D at 128 + 32 t
__host__ __device__ inline uint32_t tc5_kmaj_off(int row, int k) {  // byte offset of (row, k) in a K-major tile
  return (uint32_t)(row >> 3) * kTc5Sbo + (uint32_t)(k >> 2) * kTc5Lbo + (uint32_t)(row & 7) * 16u + (uint32_t)(k & 3) * 4u;
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo);

// ---- classifier re-layout (one small launch per overlap call) --------------
template <typename TIN>
__global__ void overlap_prep_kernel(const TIN* __restrict__ clf, OverlapParams P, OverlapFast F,
                                    float* __restrict__ small_w, float* __restrict__ wc, unsigned swz_mask = 0u) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  for (int i = tid; i < F.small_w_total; i += nth) small_w[i] = 0.f;
  for (int64_t i = tid; i < F.wc_elems; i += nth) wc[i] = 0.f;
  // launched as ONE block (the buffers are < 150 KB), so a block barrier orders
  // the zero fill before the scatter
  __syncthreads();
  for (int n = 0; n < P.L; ++n) {
    const int Bl = P.B[n], Br = P.B[n + 1];
    const TIN* W = clf + P.clf_off[n];
    if (n == P.c) {
      const int Br4 = r4(Br);
      for (int i = tid; i < 2 * P.S * Bl * Br; i += nth) {
        const int bp = i % Br, b = (i / Br) % Bl, l = (i / (Br * Bl)) % P.S, s = i / (Br * Bl * P.S);
        wc[((int64_t)(s * Bl + b) * Br4 + bp) * F.Spad + l] = (float)W[i];
      }
    } else if (n < P.c) {
      const int ld = r4(Br);
      for (int i = tid; i < 2 * Bl * Br; i += nth) {
        const int bp = i % Br, sb = i / Br;  // sb = s*Bl + b
        // swz_mask bit n: this site is a tensor-core fragment operand (rows of 32): bank swizzle, see swz<32>
        const int col = ((swz_mask >> n) & 1u) ? (bp ^ ((sb & 3) << 3)) : bp;
        small_w[F.w_off[n] + (int64_t)sb * ld + col] = (float)W[i];
      }
    } else {
      const int ld = r4(Bl);  // transposed: [s][B'][B]
      for (int i = tid; i < 2 * Bl * Br; i += nth) {
        const int bp = i % Br, b = (i / Br) % Bl, s = i / (Br * Bl);
        small_w[F.w_off[n] + (int64_t)(s * Br + bp) * ld + b] = (float)W[i];
      }
    }
  }
  if (F.w4u_off >= 0) {
    // site 4 (2, 1, 16, 32) as the tcgen05 B operand [n = B'][k = (s,B)], split into TF32 hi / lo
    const TIN* W = clf + P.clf_off[4];
    for (int i = tid; i < 2 * 16 * 32; i += nth) {
      const int bp = i & 31, k = i >> 5;  // k = s * 16 + B
      uint32_t hi, lo;
      split_tf32((float)W[i], hi, lo);
      const int o = (int)(tc5_kmaj_off(bp, k) >> 2);
      small_w[F.w4u_off + o] = __uint_as_float(hi);
      small_w[F.w4u_off + kTc5TileFloats + o] = __uint_as_float(lo);
    }
  }
}

// C[m][n] = sum_{k<K} Xt[k*ldx + m] * Y[k*ldy + n], m < M, n < N.  4x4 tiles.
// Reads may touch the (finite or not) padding up to r4(M), r4(N): such values
// only reach accumulators that are never stored.
__device__ __forceinline__ void mm_kmajor(float* __restrict__ C, int ldc, const float* __restrict__ Xt,
                                          int ldx, const float* __restrict__ Y, int ldy, int M, int N,
                                          int K, int lane) {
  const int tm = (M + 3) >> 2, tn = (N + 3) >> 2, nt = tm * tn;
  // tm is a power of two for the natural bond profiles: shift instead of divide
  const bool p2 = (tm & (tm - 1)) == 0;
  const int sh = 31 - __clz(tm);
  for (int t = lane; t < nt; t += 32) {
    const int tj = p2 ? (t >> sh) : t / tm, ti = t - tj * tm;
    const float* xp = Xt + 4 * ti;
    const float* yp = Y + 4 * tj;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      const float4 x = *reinterpret_cast<const float4*>(xp + k * ldx);
      const float4 y = *reinterpret_cast<const float4*>(yp + k * ldy);
      const float xv[4] = {x.x, x.y, x.z, x.w};
      const float yv[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xv[i], yv[j], acc[i][j]);
    }
    const int m0 = 4 * ti, n0 = 4 * tj;
    if (n0 + 3 < N) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (m0 + i < M)
          *reinterpret_cast<float4*>(C + (m0 + i) * ldc + n0) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (m0 + i < M && n0 + j < N) C[(m0 + i) * ldc + n0 + j] = acc[i][j];
    }
  }
}

// The large contractions of the chain (M * N * nslice == 1024 outputs, e.g. the
// two s-slices of a 16 x 32 or one 32 x 32): exactly one 8 x 4 register tile
// per lane, 3 LDS.128 per 32 FMA, no tile loop.  Slice s uses C + s*offC,
// Xt + s*offX, Y + s*offY.  M % 8 == 0, N % 4 == 0, tiles per slice a power of two.
__device__ __forceinline__ void mm_kmajor_84(float* __restrict__ C, int ldc, int offC,
                                             const float* __restrict__ Xt, int ldx, int offX,
                                             const float* __restrict__ Y, int ldy, int offY, int M, int N,
                                             int K, int lane) {
  const int tm = M >> 3, tn = N >> 2;
  const int per = tm * tn;                  // tiles per slice (16 or 32)
  const int s = lane >> (31 - __clz(per)), t = lane & (per - 1);
  const int shm = 31 - __clz(tm);
  const int tj = t >> shm, ti = t - (tj << shm);
  const float* xp = Xt + s * offX + 8 * ti;
  const float* yp = Y + s * offY + 4 * tj;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    const float4 x0 = *reinterpret_cast<const float4*>(xp + k * ldx);
    const float4 x1 = *reinterpret_cast<const float4*>(xp + k * ldx + 4);
    const float4 y = *reinterpret_cast<const float4*>(yp + k * ldy);
    const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
    const float yv[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xv[i], yv[j], acc[i][j]);
  }
  float* cp = C + s * offC + (8 * ti) * ldc + 4 * tj;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    *reinterpret_cast<float4*>(cp + i * ldc) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
}

// nslice contractions of the same shape; picks the 8x4 single-tile form when it applies
__device__ __forceinline__ void mm_slices(float* __restrict__ C, int ldc, int offC, const float* __restrict__ Xt,
                                          int ldx, int offX, const float* __restrict__ Y, int ldy, int offY,
                                          int M, int N, int K, int nslice, int lane) {
  if (M * N * nslice == 1024 && (M & 7) == 0 && (N & 3) == 0) {
    mm_kmajor_84(C, ldc, offC, Xt, ldx, offX, Y, ldy, offY, M, N, K, lane);
  } else if (M * N == 1024 && (M & 7) == 0 && (N & 3) == 0) {
    // 1,024 outputs per slice (last-site-hairy classifiers: 32 x 32 environments): one 8 x 4 tile
    // per lane and slice
    for (int s = 0; s < nslice; ++s)
      mm_kmajor_84(C + s * offC, ldc, 0, Xt + s * offX, ldx, 0, Y + s * offY, ldy, 0, M, N, K, lane);
  } else {
    for (int s = 0; s < nslice; ++s)
      mm_kmajor(C + s * offC, ldc, Xt + s * offX, ldx, Y + s * offY, ldy, M, N, K, lane);
  }
}


// ---- compile-time shapes: the natural 32 / 32 profile (L = 10, c = 5, 16 labels) ----
// Same contractions, same operands and the same order of accumulation along k as
// the run-time versions above, but every shape, stride and tile count is a
// constant: each phase is one pass of <= 32 register tiles whose size is chosen
// so that all lanes work (1024 outputs: 8 x 4, 256: 4 x 2, 64: 2 x 1, fewer: 1 x 1),
// and the per-phase index arithmetic of mm_kmajor disappears.
template <int W>
__device__ __forceinline__ void ldv(float (&d)[W], const float* __restrict__ p) {
  if constexpr (W == 8) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
  } else if constexpr (W == 4) {
    const float4 a = *reinterpret_cast<const float4*>(p);
    d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w;
  } else if constexpr (W == 2) {
    const float2 a = *reinterpret_cast<const float2*>(p);
    d[0] = a.x; d[1] = a.y;
  } else {
    d[0] = *p;
  }
}
template <int W>
__device__ __forceinline__ void stv(float* __restrict__ p, const float (&v)[W]) {
  if constexpr (W == 4) *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  else if constexpr (W == 2) *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
  else p[0] = v[0];
}

// NS slices of C[m][n] = sum_k Xt[k][m] Y[k][n]  (M x N, k < K), slice s at C + s*OFFC, Xt + s*OFFX, Y + s*OFFY
// Bank swizzles of the buffers the tensor-core contractions read as fragments (lane = (g, t)
// loads element (k0 + t, c0 + g)): with rows of 32 floats the four k rows of a fragment would
// sit on the same banks, so column c of row k is stored at c ^ ((k & 3) << 3); with rows of 16,
// at c ^ (((k >> 1) & 1) << 3).  Multiples of 8: float2 / float4 accesses stay contiguous.
template <int SW>
__device__ __forceinline__ int swz(int k, int col) {
  if constexpr (SW == 32) return col ^ ((k & 3) << 3);
  else if constexpr (SW == 16) return col ^ (((k >> 1) & 1) << 3);
  else return col;
}

template <int M, int N, int K, int NS, int LDC, int OFFC, int LDX, int OFFX, int LDY, int OFFY, int CSW = 0>
__device__ __forceinline__ void mm_static(float* __restrict__ C, const float* __restrict__ Xt,
                                          const float* __restrict__ Y, int lane) {
  constexpr int OUT = M * N * NS;
  // outputs per lane ~ OUT / 32: 8x4 (>= 1024; 2,048 outputs take two passes), 4x4, 4x2, 2x2, 2x1, 1x1
  constexpr int TM = OUT >= 1024 ? 8 : (OUT >= 256 ? 4 : (OUT >= 64 ? 2 : 1));
  constexpr int TN = OUT >= 512 ? 4 : (OUT >= 128 ? 2 : 1);
  static_assert(M % TM == 0 && N % TN == 0, "tile does not divide the shape");
  constexpr int tm = M / TM, tn = N / TN, per = tm * tn, NT = per * NS;
  static_assert(NT <= 64, "more than two passes of tiles");
#pragma unroll
  for (int t0 = 0; t0 < NT; t0 += 32) {
    const int tile = t0 + lane;
    if (tile < NT) {
      const int s = tile / per, t = tile % per;
      const int tj = t / tm, ti = t % tm;
      const float* xp = Xt + s * OFFX + TM * ti;
      const float* yp = Y + s * OFFY + TN * tj;
      float acc[TM][TN];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
#pragma unroll(K <= 8 ? K : 8)
      for (int k = 0; k < K; ++k) {
        float xv[TM], yv[TN];
        ldv<TM>(xv, xp + k * LDX);
        ldv<TN>(yv, yp + k * LDY);
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(xv[i], yv[j], acc[i][j]);
      }
      float* cp = C + s * OFFC + (TM * ti) * LDC;
#pragma unroll
      for (int i = 0; i < TM; ++i) stv<TN>(cp + i * LDC + swz<CSW>(TM * ti + i, TN * tj), acc[i]);
    }
  }
}

// ---- tensor-core contractions (mma.sync m16n8k8, TF32 inputs, fp32 accumulate) ----
// The 8 x 4 FFMA tiles above read 12 floats per lane and k step: 3 LDS.128 = 12 shared-memory
// wavefronts per 32 FFMA, and the four 1,024-output contractions of the chain were bound by
// the LSU pipe (88 % of its wavefront peak, profiles/ncu_overlap_r1.csv), not by the FMA pipe.
// An MMA fragment load touches every operand element once per warp (16 LDS.32 per k step of 8
// for a 32 x 32 x 8 product).  TF32 has a 10-bit mantissa, the parity bar is 1e-5, so every
// operand is split x = hi + lo (both TF32) and three MMAs (lo.hi, hi.lo, hi.hi) give products
// to ~2^-20: "3xTF32".
// hi = x with the mantissa cut to TF32's 10 bits, lo = x - hi (exact, <= 13 significant bits).
// The tensor core reads only the upper 19 bits of a TF32 operand, so adding half an ulp of that
// format to lo's bit pattern makes the hardware's truncation a rounding to nearest: the
// representation error of x is <= 2^-21 |x| and unbiased.  (cvt.rna.tf32.f32 is emulated with
// ~5 instructions on sm_100a; two of them per operand element were 40 % of the kernel's
// instructions.)
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi)) + 0x1000u;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// NS slices of C[m][n] = sum_k Xt[k][m] Y[k][n]; Xt rows of LDX floats and Y rows of LDY floats
// are swizzled (swz<LDX>, swz<LDY>), C is written with swz<CSW>.  OFFX == 0: Xt shared by the slices.
// XROW: the first operand is stored m-major instead, X[m][k] with rows of 16 floats (K = 16, the
// MPS's own site layout) and its 16-byte quads swizzled, quad q of row m at q ^ ((m >> 1) & 3).
template <int M, int N, int K, int NS, int LDC, int OFFC, int CSW, int LDX, int OFFX, int LDY, int OFFY,
          bool XROW = false>
__device__ __forceinline__ void mma_static(float* __restrict__ C, const float* __restrict__ Xt,
                                           const float* __restrict__ Y, int lane) {
  constexpr int MT = M / 16, NTL = N / 8, KS = K / 8;
  static_assert(M % 16 == 0 && N % 8 == 0 && K % 8 == 0, "mma tile");
  const int g = lane >> 2, t = lane & 3;
  float acc[NS][MT][NTL][4];
#pragma unroll
  for (int s = 0; s < NS; ++s)
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NTL; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[s][mt][nt][i] = 0.f;
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    const int k0 = 8 * ks + t, k1 = k0 + 4;
    uint32_t ah[MT][4], al[MT][4];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      if (OFFX != 0 || s == 0) {
        const float* x0 = Xt + s * OFFX + k0 * LDX;
        const float* x1 = Xt + s * OFFX + k1 * LDX;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const int m = 16 * mt + g;
          if constexpr (XROW) {
            const float* xb = Xt + s * OFFX;
            const int r0 = (m >> 1) & 3, r1 = ((m + 8) >> 1) & 3;
            split_tf32(xb[m * 16 + ((((k0 >> 2) ^ r0) << 2) | (k0 & 3))], ah[mt][0], al[mt][0]);
            split_tf32(xb[(m + 8) * 16 + ((((k0 >> 2) ^ r1) << 2) | (k0 & 3))], ah[mt][1], al[mt][1]);
            split_tf32(xb[m * 16 + ((((k1 >> 2) ^ r0) << 2) | (k1 & 3))], ah[mt][2], al[mt][2]);
            split_tf32(xb[(m + 8) * 16 + ((((k1 >> 2) ^ r1) << 2) | (k1 & 3))], ah[mt][3], al[mt][3]);
          } else {
            split_tf32(x0[swz<LDX>(k0, m)], ah[mt][0], al[mt][0]);
            split_tf32(x0[swz<LDX>(k0, m + 8)], ah[mt][1], al[mt][1]);
            split_tf32(x1[swz<LDX>(k1, m)], ah[mt][2], al[mt][2]);
            split_tf32(x1[swz<LDX>(k1, m + 8)], ah[mt][3], al[mt][3]);
          }
        }
      }
      const float* y0 = Y + s * OFFY + k0 * LDY;
      const float* y1 = Y + s * OFFY + k1 * LDY;
#pragma unroll
      for (int nt = 0; nt < NTL; ++nt) {
        const int n = 8 * nt + g;
        uint32_t bh[2], bl[2];
        split_tf32(y0[swz<LDY>(k0, n)], bh[0], bl[0]);
        split_tf32(y1[swz<LDY>(k1, n)], bh[1], bl[1]);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          mma_tf32(acc[s][mt][nt], al[mt], bh);
          mma_tf32(acc[s][mt][nt], ah[mt], bl);
          mma_tf32(acc[s][mt][nt], ah[mt], bh);
        }
      }
    }
  }
#pragma unroll
  for (int s = 0; s < NS; ++s)
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NTL; ++nt) {
        const int r = 16 * mt + g, col = 8 * nt + 2 * t;
        float* c0 = C + s * OFFC + r * LDC + swz<CSW>(r, col);
        float* c1 = C + s * OFFC + (r + 8) * LDC + swz<CSW>(r + 8, col);
        if constexpr (CSW == 132) {
          // rows of LDC = 32 floats holding the slices side by side (OFFC columns apart); the 16-byte chunks
          // of row r rotated by r & 7 so that lane = row reads its row with conflict-free 128-bit loads
          c0 = C + r * LDC + ((s * OFFC + col) ^ ((r & 7) << 2));
          c1 = C + (r + 8) * LDC + ((s * OFFC + col) ^ (((r + 8) & 7) << 2));
        }
        *reinterpret_cast<float2*>(c0) = make_float2(acc[s][mt][nt][0], acc[s][mt][nt][1]);
        *reinterpret_cast<float2*>(c1) = make_float2(acc[s][mt][nt][2], acc[s][mt][nt][3]);
      }
}

// Two independent contractions of the SAME static shape, one per half-warp.  Left-sweep step n and
// right-sweep step 9 - n of the natural profile are mirror images (same M, N, K and strides), and
// each is a chain of short dependent phases: running them side by side halves the number of
// phases (16 -> 8 for sites 0-3 / 9-6).  CSW1: output swizzle of the second problem.
template <int M, int N, int K, int NS, int LDC, int OFFC, int LDX, int OFFX, int LDY, int OFFY, int CSW1 = 0,
          int CSW0 = 0>
__device__ __forceinline__ void mm_static_pair(float* __restrict__ C0, const float* __restrict__ X0,
                                               const float* __restrict__ Y0, float* __restrict__ C1,
                                               const float* __restrict__ X1, const float* __restrict__ Y1,
                                               int lane) {
  constexpr int OUT = M * N * NS;
  constexpr int TM = OUT >= 256 ? 4 : (OUT >= 64 ? 2 : 1);
  constexpr int TN = TM;
  static_assert(M % TM == 0 && N % TN == 0, "tile does not divide the shape");
  constexpr int tm = M / TM, tn = N / TN, per = tm * tn, NT = per * NS;
  static_assert(NT <= 16, "more tiles than lanes of a half-warp");
  const int h = lane >> 4, l = lane & 15;
  if (NT < 16 && l >= NT) return;
  float* C = h ? C1 : C0;
  const float* Xt = h ? X1 : X0;
  const float* Y = h ? Y1 : Y0;
  const int s = l / per, t = l % per;
  const int tj = t / tm, ti = t % tm;
  const float* xp = Xt + s * OFFX + TM * ti;
  const float* yp = Y + s * OFFY + TN * tj;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
#pragma unroll(K <= 8 ? K : 8)
  for (int k = 0; k < K; ++k) {
    float xv[TM], yv[TN];
    ldv<TM>(xv, xp + k * LDX);
    ldv<TN>(yv, yp + k * LDY);
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(xv[i], yv[j], acc[i][j]);
  }
  float* cp = C + s * OFFC + (TM * ti) * LDC;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int col = h ? swz<CSW1>(TM * ti + i, TN * tj) : swz<CSW0>(TM * ti + i, TN * tj);
    stv<TN>(cp + i * LDC + col, acc[i]);
  }
}

// left step n (sites 0..3) and right step 9 - n side by side; R ping-pongs between Rcur and Rnext
// SWAPL: the left problem's result transposed, E'[a'][B'] instead of Et'[B'][a'] (same shape when ar == Br)
template <int NSITE, int CSW1 = 0, int CSW0 = 0, bool SWAPL = false>
__device__ __forceinline__ void nat_pair_step(float* E, float* Tb, const float* W, const float* A, const float* Rcur,
                                              float* Rnext, const float* Wt, const float* At, int lane) {
  constexpr int al = 1 << NSITE, ar = 2 << NSITE, Bl = al, Br = ar;
  constexpr int ldE = (al + 3) & ~3, ldW = (Br + 3) & ~3, ldA = (ar + 3) & ~3, ldT = ldW;
  // T[(s,a)][B'] = sum_B Et[B][a] W[s][B][B']   ||   T'[s][B'][a] = sum_a' R[a'][B'] At[s][a'][a]
  mm_static_pair<al, Br, Bl, 2, ldT, al * ldT, ldE, 0, ldW, Bl * ldW>(Tb, E, W, Tb + 1024, Rcur, At, lane);
  __syncwarp();
  // Et'[B'][a'] = sum_(s,a) T[(s,a)][B'] A[(s,a)][a']   ||   R'[a][B] = sum_(s,B') T'[(s,B')][a] Wt[(s,B')][B]
  if constexpr (SWAPL) {
    static_assert(ar == Br && ldA == ldT, "transposed left result needs a square step");
    mm_static_pair<Br, ar, 2 * al, 1, ldA, 0, ldT, 0, ldA, 0, CSW1, CSW0>(E, A, Tb, Rnext, Tb + 1024, Wt, lane);
  } else {
    mm_static_pair<Br, ar, 2 * al, 1, ldA, 0, ldT, 0, ldA, 0, CSW1, CSW0>(E, Tb, A, Rnext, Tb + 1024, Wt, lane);
  }
  __syncwarp();
}

// ---- tcgen05 helpers (MODE 6) ----
__device__ __forceinline__ uint64_t tc5_smem_desc(uint32_t saddr) {  // K-major, no swizzle, descriptor version 1
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((kTc5Lbo >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((kTc5Sbo >> 4) & 0x3fffu) << 32) | ((uint64_t)1 << 46);
}
// D[tmem] (+)= A[tmem] * B[smem descriptor], kind::tf32, issued by one thread for the CTA
__device__ __forceinline__ void tc5_mma(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tc5_st16(uint32_t taddr, const uint32_t (&v)[16]) {  // lane = TMEM lane, 16 columns
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tc5_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}

template <int NSITE>
__device__ __forceinline__ void nat_left_step(float* E, float* Tb, const float* W, const float* A, int lane) {
  constexpr int al = 1 << NSITE, ar = 2 << NSITE, Bl = al, Br = ar;
  constexpr int ldE = (al + 3) & ~3, ldW = (Br + 3) & ~3, ldA = (ar + 3) & ~3, ldT = ldW;
  // T[(s,a)][B'] = sum_B Et[B][a] W[s][B][B']
  mm_static<al, Br, Bl, 2, ldT, al * ldT, ldE, 0, ldW, Bl * ldW>(Tb, E, W, lane);
  __syncwarp();
  if constexpr (NSITE + 1 < 5) {
    // Et'[B'][a'] = sum_(s,a) T[(s,a)][B'] A[(s,a)][a']
    mm_static<Br, ar, 2 * al, 1, ldA, 0, ldT, 0, ldA, 0>(E, Tb, A, lane);
  } else {
    // EL[a'][B'] = sum_(s,a) A[(s,a)][a'] T[(s,a)][B']
    mm_static<ar, Br, 2 * al, 1, ldT, 0, ldA, 0, ldT, 0>(E, A, Tb, lane);
  }
  __syncwarp();
}

template <int NSITE, int CSW = 0>
__device__ __forceinline__ void nat_right_step(float* R, float* R2, float* Tb, const float* Wt, const float* At,
                                               int lane) {
  constexpr int al = 1 << (10 - NSITE), ar = 1 << (9 - NSITE), Bl = al, Br = ar;
  constexpr int ldR = (Br + 3) & ~3, ldAt = (al + 3) & ~3, ldWt = (Bl + 3) & ~3, ldTp = ldAt;
  // T'[s][B'][a] = sum_a' R[a'][B'] At[s][a'][a]
  mm_static<Br, al, ar, 2, ldTp, Br * ldTp, ldR, 0, ldAt, ar * ldAt>(Tb, R, At, lane);
  __syncwarp();
  // R'[a][B] = sum_(s,B') T'[(s,B')][a] Wt[(s,B')][B]
  mm_static<al, Bl, 2 * Br, 1, ldWt, 0, ldTp, 0, ldWt, 0, CSW>(R2, Tb, Wt, lane);
  __syncwarp();
}

// left step n (0..8) of the LAST-site-hairy natural-32 profile: image bonds natural, classifier
// bonds [1,2,4,8,16,32,32,32,32,32,1] (fMPO.compress_one_site, fMPO.py:240-336)
template <int NSITE>
__device__ __forceinline__ void last_left_step(float* E, float* Tb, const float* W, const float* A, int lane) {
  constexpr int al = NSITE <= 5 ? 1 << NSITE : 1 << (10 - NSITE);
  constexpr int ar = NSITE < 5 ? 2 << NSITE : 1 << (9 - NSITE);
  constexpr int Bl = NSITE <= 5 ? 1 << NSITE : 32, Br = NSITE < 5 ? 2 << NSITE : 32;
  constexpr int ldE = (al + 3) & ~3, ldW = (Br + 3) & ~3, ldA = (ar + 3) & ~3, ldT = ldW;
  // T[(s,a)][B'] = sum_B Et[B][a] W[s][B][B']
  mm_static<al, Br, Bl, 2, ldT, al * ldT, ldE, 0, ldW, Bl * ldW>(Tb, E, W, lane);
  __syncwarp();
  if constexpr (NSITE + 1 < 9) {
    // Et'[B'][a'] = sum_(s,a) T[(s,a)][B'] A[(s,a)][a']
    mm_static<Br, ar, 2 * al, 1, ldA, 0, ldT, 0, ldA, 0>(E, Tb, A, lane);
  } else {
    // EL[a'][B'] = sum_(s,a) A[(s,a)][a'] T[(s,a)][B']
    mm_static<ar, Br, 2 * al, 1, ldT, 0, ldA, 0, ldT, 0>(E, A, Tb, lane);
  }
  __syncwarp();
}

inline bool overlap_is_last_natural(const OverlapParams& P) {
  if (P.L != 10 || P.c != 9 || P.S != 16) return false;
  for (int n = 0; n <= 10; ++n) {
    const int a = n <= 5 ? (1 << n) : (1 << (10 - n));
    const int b = n == 10 ? 1 : (n <= 5 ? (1 << n) : 32);
    if (P.B[n] != b || P.a[n] > a || P.a[n] < 1) return false;
  }
  return true;
}

inline bool overlap_is_natural(const OverlapParams& P) {
  if (P.L != 10 || P.c != 5 || P.S != 16) return false;
  for (int n = 0; n <= 10; ++n) {
    const int b = n <= 5 ? (1 << n) : (1 << (10 - n));
    if (P.a[n] != b || P.B[n] != b) return false;
  }
  return true;
}
// natural-32 classifier, image bonds anywhere below the natural profile (D_encode < 32)
inline bool overlap_is_padded_natural(const OverlapParams& P) {
  if (P.L != 10 || P.c != 5 || P.S != 16) return false;
  bool smaller = false;
  for (int n = 0; n <= 10; ++n) {
    const int b = n <= 5 ? (1 << n) : (1 << (10 - n));
    if (P.B[n] != b || P.a[n] > b || P.a[n] < 1) return false;
    smaller = smaller || P.a[n] < b;
  }
  return smaller;
}

// MODE 0: run-time shapes; 1: natural 32/32 profile, compile-time shapes, FFMA; 2: the same with
// the GEMM-shaped contractions on the tensor cores (3xTF32); 3: as 2 for image MPSs of any smaller
// bond dimension (D_encode < 32) against the natural-32 classifier: the sites are staged
// zero-padded to the natural profile (the padding is written once per warp), so a truncated
// encoding costs what the full one does instead of the run-time-shaped kernel's 2.5x;
// 4: last-site-hairy natural-32 classifier (c = 9), compile-time shapes, FFMA, images of any
// D_encode <= 32 staged zero-padded; 5: the same with the contractions of sites 4-6 (70 % of the
// MACs) on the tensor cores
template <int MODE>
__global__ void __launch_bounds__(256, 1)
overlap_fast_kernel(const __grid_constant__ OverlapParams P, const __grid_constant__ OverlapFast F,
                    const float* __restrict__ small_w, const float* __restrict__ wc) {
  constexpr bool TC5 = MODE == 6;
  constexpr bool NAT = (MODE >= 1 && MODE <= 3) || TC5, MMA = MODE == 2 || MODE == 3 || TC5, PAD = MODE == 3;
  constexpr bool LAST = MODE == 4 || MODE == 5, LMMA = MODE == 5;
  extern __shared__ __align__(128) unsigned char smem_ov[];
  float* sW = reinterpret_cast<float*>(smem_ov);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  constexpr int PER_WARP = TC5 ? kTc5PerWarp : kOvPerWarp;
  constexpr int MAXA = TC5 ? kTc5MaxA : kOvMaxA;
  float* wbase = sW + ((F.small_w_total + 31) & ~31) + (size_t)warp * PER_WARP;
  float* As = wbase;
  float* Eb = As + MAXA;
  float* Tb = Eb + kOvE;
  float* Rb0 = Tb + kOvT;
  float* Rb1 = Rb0 + kOvR;
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ __align__(8) unsigned long long mbar_tc;  // TC5: completion of the CTA's tcgen05.mma batch
  __shared__ uint32_t tmem_base_s;
  if constexpr (TC5) {
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&mbar_tc)));
    }
    if (warp == 0) {  // one CTA per SM (213 KB of shared memory): the allocation cannot contend
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                       (unsigned)__cvta_generic_to_shared(&tmem_base_s)), "n"(kTc5Cols));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }

  // ---- one bulk TMA copy of the small classifier sites into shared memory ----
  if (threadIdx.x == 0) {
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&mbar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0 && F.small_w_total > 0) {
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&mbar);
    const unsigned dst = (unsigned)__cvta_generic_to_shared(sW);
    const unsigned bytes = (unsigned)F.small_w_total * 4u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
        "l"(small_w), "r"(bytes), "r"(bar)
        : "memory");
  }
  if (F.small_w_total > 0) {
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&mbar);
    unsigned done = 0;
    for (int spin = 0; !done; ++spin) {
      if (spin > (1 << 24)) __trap();  // a lost TMA completion must not hang the GPU
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(bar)
          : "memory");
    }
  }

  uint32_t tmem_base = 0;
  if constexpr (TC5) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    tmem_base = tmem_base_s;  // written before the __syncthreads above
  }
  const int c = P.c, L = P.L, S = P.S;
  // ---- per-CTA staging table: element e of an image block -> offset in As.
  //      left sites keep [s][a][a'] (ld r4(a')), hairy and right sites are
  //      transposed to [s][a'][a] (ld r4(a)) ----
  unsigned short* tab = reinterpret_cast<unsigned short*>(sW + ((F.small_w_total + 31) & ~31) +
                                                          (size_t)nwarps * PER_WARP);
  const int stride = (int)P.mps_stride;
  for (int e = threadIdx.x; e < stride; e += blockDim.x) {
    int n = 0;
    while (n + 1 < L && e >= (int)P.mps_off[n + 1]) ++n;
    const int i = e - (int)P.mps_off[n];
    int q, ap;  // element (s, a, a') of site n: q = s*al + a, ap = a'
    if (P.mirrored) {
      // block position n holds site L-1-n with its two bond indices swapped (symmetric bond profile:
      // the positions' offsets are the sites' offsets)
      n = L - 1 - n;
      const int al_ = P.a[n], ar_ = P.a[n + 1];
      const int y = i % al_, sx = i / al_;
      q = (sx / ar_) * al_ + y;
      ap = sx % ar_;
    } else {
      q = i / P.a[n + 1];
      ap = i - q * P.a[n + 1];
    }
    const int al = P.a[n], ar = P.a[n + 1];
    int d;
    if constexpr (LAST) {
      // natural-profile destination; every site but the last keeps [s][a][a'] (ld r4(a')), the
      // hairy site 9 is [s][a'][a]
      const int AL = n <= 5 ? 1 << n : 1 << (10 - n), AR = n < 5 ? 2 << n : 1 << (9 - n);
      const int sdx = q / al, a = q - sdx * al;
      d = n < 9 ? (sdx * AL + a) * r4(AR) + ap : (sdx * AR + ap) * r4(AL) + a;
      if (LMMA && n == 4) d = (sdx * 16 + a) * 32 + swz<32>(sdx * 16 + a, ap);   // fragment operands
      if (LMMA && n == 5) d = (sdx * 32 + a) * 16 + swz<16>(sdx * 32 + a, ap);
    } else if constexpr (PAD) {
      // destination in the NATURAL profile's layout (F.a_off is the natural one), with the
      // fragment swizzles of sites 4 and 5
      const int AL = n <= 5 ? 1 << n : 1 << (10 - n), AR = n < 5 ? 2 << n : 1 << (9 - n);
      const int sdx = q / al, a = q - sdx * al;
      if (n < 4) d = (sdx * AL + a) * r4(AR) + ap;
      else if (n == 4) d = (sdx * 16 + a) * 32 + swz<32>(sdx * 16 + a, ap);
      else if (n == 5) d = sdx * 512 + a * 16 + ((((ap >> 2) ^ ((a >> 1) & 3)) << 2) | (ap & 3));
      else d = (sdx * AR + ap) * r4(AL) + a;
    } else if (n < c) {
      d = q * r4(ar) + ap;
      if (MMA && n == 4) d = q * 32 + swz<32>(q, ap);  // fragment operand of the site-4 contraction
    } else {
      const int sdx = q / al, a = q - sdx * al;
      d = (sdx * ar + ap) * r4(al) + a;
      // natural profile, tensor-core path: site 5 keeps its [s][a][a'] layout (the m-major operand of G),
      // its 16-byte quads swizzled
      if (MMA && n == c) d = sdx * 512 + a * 16 + ((((ap >> 2) ^ ((a >> 1) & 3)) << 2) | (ap & 3));
    }
    tab[e] = (unsigned short)(F.a_off[n] + d);
  }
  __syncthreads();
  // vector path: whole image block = `stride` floats, 16-byte aligned, <= kOvPre*128
  const bool vec = (stride & 3) == 0 && stride <= kOvPre * 128 &&
                   ((reinterpret_cast<uintptr_t>(P.mps) & 15) == 0);
  float4 pre[kOvPre];
  auto prefetch = [&](int64_t im) {
    const float4* src = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(P.mps) + (size_t)im * stride);
#pragma unroll
    for (int j = 0; j < kOvPre; ++j) {
      const int v = lane + 32 * j;
      if (4 * v < stride) pre[j] = __ldcs(src + v);
    }
  };
  // the same in four bursts (natural profile): 22 back-to-back 512-byte warp loads from eight warps
  // ran into the load/store unit's queue limit (lg_throttle, a fifth of the samples)
  auto prefetch_part = [&](int64_t im, int part) {
    const float4* src = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(P.mps) + (size_t)im * stride);
#pragma unroll
    for (int j = 0; j < kOvPre; ++j) {
      const int v = lane + 32 * j;
      if ((j * 4) / kOvPre == part && 4 * v < stride) pre[j] = __ldcs(src + v);
    }
  };
  if constexpr (PAD || LAST) {
    for (int i = lane; i < MAXA; i += 32) As[i] = 0.f;  // the padding stays zero: the scatter never touches it
    __syncwarp();
  }
  [[maybe_unused]] unsigned tc_phase = 0u;  // TC5: parity of the CTA iteration's tcgen05 completion
  if (vec && (int64_t)blockIdx.x * nwarps + warp < P.N) prefetch((int64_t)blockIdx.x * nwarps + warp);
  // the trip count is per CTA (the tensor-core label sums below are CTA-wide); a warp whose image
  // index runs past N only takes part in the barriers
  for (int64_t base = (int64_t)blockIdx.x * nwarps; base < P.N; base += (int64_t)gridDim.x * nwarps) {
    const int64_t img = base + warp;
    const bool valid = img < P.N;
    if (valid) {
    const float* mps = reinterpret_cast<const float*>(P.mps) + (size_t)img * P.mps_stride;
    if (vec) {
#pragma unroll
      for (int j = 0; j < kOvPre; ++j) {
        const int v = lane + 32 * j;
        if (MMA && !PAD && v >= 341 && v <= 596) {
          // site 5 keeps its [s][a][a'] layout (the m-major operand of G), quads swizzled.  (The places
          // are computed, not read from the staging table: a table load in front of every store measured
          // 525 vs 519 us - the kernel is bound by latency, not by its instruction count.)
          const int i = 4 * v - 1364;
          const int a = (i >> 4) & 31, q = (i >> 2) & 3;
          *reinterpret_cast<float4*>(As + 1368 + (i & ~15) + ((q ^ ((a >> 1) & 3)) << 2)) = pre[j];
        } else if (NAT && !PAD && v >= 1 && v <= 340) {
          // sites 1..4 of the natural profile keep their layout ([s][a][a'], no row padding):
          // a straight 128-bit copy, 4 floats further on (site 0 is padded from 4 to 8 floats);
          // site 4 (v >= 85) is a fragment operand of the tensor-core path: swizzled rows
          int dst = 4 * v + 4;
          if (MMA && v >= 85) {
            const int i = 4 * v - 340;
            dst = 344 + (i & ~31) + swz<32>(i >> 5, i & 31);
          }
          *reinterpret_cast<float4*>(As + dst) = pre[j];
        } else if (4 * v < stride) {
          const uint2 t = *reinterpret_cast<const uint2*>(tab + 4 * v);
          As[t.x & 0xffff] = pre[j].x;
          As[t.x >> 16] = pre[j].y;
          As[t.y & 0xffff] = pre[j].z;
          As[t.y >> 16] = pre[j].w;
        }
      }
      const int64_t nxt = img + (int64_t)gridDim.x * nwarps;
      if (nxt < P.N) {  // in flight during this image's contractions
        if constexpr (NAT) prefetch_part(nxt, 0);
        else prefetch(nxt);
      }
    } else {
      for (int e = lane; e < stride; e += 32) As[tab[e]] = __ldg(mps + e);
    }
    __syncwarp();

    if constexpr (LAST) {
      // ---- last-site-hairy natural-32 profile: nine left steps with constant shapes ----
      float* E = Eb;
      if (lane == 0) E[0] = 1.f;
      __syncwarp();
      last_left_step<0>(E, Tb, sW + F.w_off[0], As + F.a_off[0], lane);
      last_left_step<1>(E, Tb, sW + F.w_off[1], As + F.a_off[1], lane);
      last_left_step<2>(E, Tb, sW + F.w_off[2], As + F.a_off[2], lane);
      last_left_step<3>(E, Tb, sW + F.w_off[3], As + F.a_off[3], lane);
      if constexpr (LMMA) {
        // site 4: T (FFMA, K = 16) written swizzled; Et'[B'][a'] = sum_(s,a) T[(s,a)][B'] A4[(s,a)][a']
        mm_static<16, 32, 16, 2, 32, 512, 16, 0, 32, 512, 32>(Tb, E, sW + F.w_off[4], lane);
        __syncwarp();
        mma_static<32, 32, 32, 1, 32, 0, 32, 32, 0, 32, 0>(E, Tb, As + F.a_off[4], lane);
        __syncwarp();
        // site 5: T[(s,a)][B'] = sum_B Et[B][a] W5[s][B][B'] (2 x 32x32x32), then Et' (32 x 16, K = 64)
        mma_static<32, 32, 32, 2, 32, 1024, 32, 32, 0, 32, 1024>(Tb, E, sW + F.w_off[5], lane);
        __syncwarp();
        mma_static<32, 16, 64, 1, 16, 0, 16, 32, 0, 16, 0>(E, Tb, As + F.a_off[5], lane);
        __syncwarp();
        // site 6: T (2 x 16x32x32), then Et' (32 x 8, K = 32; rows of 8 floats need no swizzle)
        mma_static<16, 32, 32, 2, 32, 512, 32, 16, 0, 32, 1024>(Tb, E, sW + F.w_off[6], lane);
        __syncwarp();
        mma_static<32, 8, 32, 1, 8, 0, 0, 32, 0, 8, 0>(E, Tb, As + F.a_off[6], lane);
        __syncwarp();
      } else {
        last_left_step<4>(E, Tb, sW + F.w_off[4], As + F.a_off[4], lane);
        last_left_step<5>(E, Tb, sW + F.w_off[5], As + F.a_off[5], lane);
        last_left_step<6>(E, Tb, sW + F.w_off[6], As + F.a_off[6], lane);
      }
      last_left_step<7>(E, Tb, sW + F.w_off[7], As + F.a_off[7], lane);
      last_left_step<8>(E, Tb, sW + F.w_off[8], As + F.a_off[8], lane);
      // hairy site 9: a = 2, a' = 1, B = 32, B' = 1: H[s][B] = sum_a EL[a][B] A9[s][a];
      // out[l] = sum_(s,B) Wc[(s,B)][l] H[s][B]   (Wc rows of r4(1) * 16 floats); lane = B
      const float* A9 = As + F.a_off[9];  // [s][a' = 0][a], ld 4
      float acc[16];
#pragma unroll
      for (int l = 0; l < 16; ++l) acc[l] = 0.f;
#pragma unroll
      for (int sdx = 0; sdx < 2; ++sdx) {
        const float h = fmaf(E[lane], A9[sdx * 4], E[32 + lane] * A9[sdx * 4 + 1]);
        const float4* wr = reinterpret_cast<const float4*>(wc + (size_t)(sdx * 32 + lane) * 64);
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const float4 w4 = __ldg(wr + q4);
          acc[4 * q4] = fmaf(w4.x, h, acc[4 * q4]);
          acc[4 * q4 + 1] = fmaf(w4.y, h, acc[4 * q4 + 1]);
          acc[4 * q4 + 2] = fmaf(w4.z, h, acc[4 * q4 + 2]);
          acc[4 * q4 + 3] = fmaf(w4.w, h, acc[4 * q4 + 3]);
        }
      }
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
        for (int l = 0; l < 16; ++l) acc[l] += __shfl_xor_sync(FULL, acc[l], o);
      }
      float v = 0.f;
#pragma unroll
      for (int l = 0; l < 16; ++l) v = lane == l ? acc[l] : v;
      if (lane < 16) reinterpret_cast<float*>(P.overlaps)[(size_t)img * 16 + lane] = P.signed_out ? v : fabsf(v);
      v = fabsf(v);
      if (P.argmax) {
        float best = lane < 16 ? v : -1.f;
        int besti = lane;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) {
          const float ob = __shfl_xor_sync(FULL, best, o);
          const int oi = __shfl_xor_sync(FULL, besti, o);
          if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
        }
        if (lane == 0) P.argmax[img] = besti;
      }
    } else if constexpr (NAT) {
      // ---- natural 32 / 32 profile: every shape a constant ----
      float* E = Eb;
      if (lane == 0) { E[0] = 1.f; Rb0[0] = 1.f; }
      __syncwarp();
      const int64_t nxt_img = img + (int64_t)gridDim.x * nwarps;
      if constexpr (TC5) {
        // sites 0..3 and 9..6 side by side; the last left step leaves E4 a-major, E4[a'][B'] (rows of 16)
        nat_pair_step<0>(E, Tb, sW + F.w_off[0], As + F.a_off[0], Rb0, Rb1, sW + F.w_off[9], As + F.a_off[9], lane);
        nat_pair_step<1>(E, Tb, sW + F.w_off[1], As + F.a_off[1], Rb1, Rb0, sW + F.w_off[8], As + F.a_off[8], lane);
        if (vec && nxt_img < P.N) prefetch_part(nxt_img, 1);
        nat_pair_step<2>(E, Tb, sW + F.w_off[2], As + F.a_off[2], Rb0, Rb1, sW + F.w_off[7], As + F.a_off[7], lane);
        nat_pair_step<3, 16, 16, true>(E, Tb, sW + F.w_off[3], As + F.a_off[3], Rb1, Rb0, sW + F.w_off[6],
                                       As + F.a_off[6], lane);
        if (vec && nxt_img < P.N) prefetch_part(nxt_img, 2);
        // image first: Y[a'][(s,B)] = sum_a A4[(s,a)][a'] E4[a][B]  (mma.sync, per image) ...
        mma_static<32, 16, 16, 2, 32, 16, 132, 32, 512, 16, 0>(Tb, As + F.a_off[4], E, lane);
        __syncwarp();
        // ... then the classifier site as the SHARED operand of one tcgen05 product per CTA:
        // row a' = lane of Y, split into TF32 hi / lo, to tensor memory (A operand: lane = row, column = k)
        {
          const float* yr = Tb + lane * 32;
          const uint32_t ta = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 64u * (uint32_t)(warp >> 2);
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const int j = 4 * half + jj;
              const float4 v = *reinterpret_cast<const float4*>(yr + ((j ^ (lane & 7)) << 2));
              split_tf32(v.x, hi[4 * jj], lo[4 * jj]);
              split_tf32(v.y, hi[4 * jj + 1], lo[4 * jj + 1]);
              split_tf32(v.z, hi[4 * jj + 2], lo[4 * jj + 2]);
              split_tf32(v.w, hi[4 * jj + 3], lo[4 * jj + 3]);
            }
            tc5_st16(ta + 16u * half, hi);
            tc5_st16(ta + 32u + 16u * half, lo);
          }
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        {
          // the valid warps of this CTA iteration (warp 0 always is one of them) meet; thread 0 issues
          const int nvalid = (int)(P.N - base < (int64_t)nwarps ? P.N - base : (int64_t)nwarps);
          asm volatile("bar.sync 1, %0;" ::"r"(32 * nvalid) : "memory");
          if (threadIdx.x == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // D = f32, A = B = TF32, K-major, N = 32, M = 128
            constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
            const uint32_t bH = (uint32_t)__cvta_generic_to_shared(sW + F.w4u_off);
            const uint32_t bL = bH + 4u * kTc5TileFloats;
            const int tiles = nvalid > 4 ? 2 : 1;
            for (int tl = 0; tl < tiles; ++tl) {
              const uint32_t dcol = tmem_base + 128u + 32u * tl, ah = tmem_base + 64u * tl, al = ah + 32u;
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                const uint32_t o = (uint32_t)ks * 2u * kTc5Lbo;  // 8 tf32 = two 16-byte chunks along K
                tc5_mma(dcol, al + 8u * ks, tc5_smem_desc(bH + o), idesc, ks > 0);
                tc5_mma(dcol, ah + 8u * ks, tc5_smem_desc(bL + o), idesc, 1u);
                tc5_mma(dcol, ah + 8u * ks, tc5_smem_desc(bH + o), idesc, 1u);
              }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                             (unsigned)__cvta_generic_to_shared(&mbar_tc))
                         : "memory");
          }
          __syncwarp();
        }
      } else if constexpr (MMA) {
        // sites 0..3 and 9..6 side by side (one half-warp each), then site 4
        nat_pair_step<0>(E, Tb, sW + F.w_off[0], As + F.a_off[0], Rb0, Rb1, sW + F.w_off[9], As + F.a_off[9], lane);
        nat_pair_step<1>(E, Tb, sW + F.w_off[1], As + F.a_off[1], Rb1, Rb0, sW + F.w_off[8], As + F.a_off[8], lane);
        if (vec && nxt_img < P.N) prefetch_part(nxt_img, 1);
        nat_pair_step<2>(E, Tb, sW + F.w_off[2], As + F.a_off[2], Rb0, Rb1, sW + F.w_off[7], As + F.a_off[7], lane);
        nat_pair_step<3, 16, 16>(E, Tb, sW + F.w_off[3], As + F.a_off[3], Rb1, Rb0, sW + F.w_off[6],
                                 As + F.a_off[6], lane);
        if (vec && nxt_img < P.N) prefetch_part(nxt_img, 2);
        // T[(s,a)][B'] = sum_B E4[B][a] W4[s][B][B'], then EL[a'][B'] = sum_(s,a) A4[(s,a)][a'] T[(s,a)][B']
        mma_static<16, 32, 16, 2, 32, 512, 32, 16, 0, 32, 512>(Tb, E, sW + F.w_off[4], lane);
        __syncwarp();
        mma_static<32, 32, 32, 1, 32, 0, 32, 32, 0, 32, 0>(E, As + F.a_off[4], Tb, lane);
        __syncwarp();
      } else {
        nat_left_step<0>(E, Tb, sW + F.w_off[0], As + F.a_off[0], lane);
        nat_left_step<1>(E, Tb, sW + F.w_off[1], As + F.a_off[1], lane);
        nat_left_step<2>(E, Tb, sW + F.w_off[2], As + F.a_off[2], lane);
        nat_left_step<3>(E, Tb, sW + F.w_off[3], As + F.a_off[3], lane);
        if (vec && nxt_img < P.N) prefetch_part(nxt_img, 1);
        nat_left_step<4>(E, Tb, sW + F.w_off[4], As + F.a_off[4], lane);
        if (vec && nxt_img < P.N) prefetch_part(nxt_img, 2);
        nat_right_step<9>(Rb0, Rb1, Tb, sW + F.w_off[9], As + F.a_off[9], lane);
        nat_right_step<8>(Rb1, Rb0, Tb, sW + F.w_off[8], As + F.a_off[8], lane);
        nat_right_step<7>(Rb0, Rb1, Tb, sW + F.w_off[7], As + F.a_off[7], lane);
        nat_right_step<6>(Rb1, Rb0, Tb, sW + F.w_off[6], As + F.a_off[6], lane);
      }
      if (vec && nxt_img < P.N) prefetch_part(nxt_img, 3);
      float* R = Rb0;
      // hairy site: al = 32, ar = 16, Bl = 32, Br = 16
      const float* At = As + F.a_off[5];  // [s][a'][a]
      float* G = Tb;                      // [s][a][B']
      float* H = Tb + 1024;               // [s][B][B']
      if constexpr (MMA) {
        mma_static<32, 16, 16, 2, 16, 512, 16, 16, 512, 16, 0, true>(G, At, R, lane);  // independent of E5
        __syncwarp();
        if constexpr (TC5) {
          // E5[a'][B'] of this image out of tensor memory: lane = a', rows of 32 floats, swz<32>
          const unsigned bar = (unsigned)__cvta_generic_to_shared(&mbar_tc);
          unsigned done = 0;
          for (int spin = 0; !done; ++spin) {
            if (spin > (1 << 24)) __trap();  // a lost completion must not hang the GPU
            asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(bar), "r"(tc_phase)
                : "memory");
          }
          tc_phase ^= 1u;
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t td = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 128u + 32u * (uint32_t)(warp >> 2);
          float* er = E + lane * 32;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t d[16];
            tc5_ld16(td + 16u * half, d);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const int j = 4 * half + jj;
              *reinterpret_cast<float4*>(er + ((j ^ ((lane & 3) << 1)) << 2)) =
                  make_float4(__uint_as_float(d[4 * jj]), __uint_as_float(d[4 * jj + 1]), __uint_as_float(d[4 * jj + 2]),
                              __uint_as_float(d[4 * jj + 3]));
            }
          }
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
        }
        mma_static<32, 16, 32, 2, 16, 512, 0, 32, 0, 16, 512>(H, E, G, lane);
      } else {
        mm_static<32, 16, 16, 2, 16, 32 * 16, 32, 16 * 32, 16, 0>(G, At, R, lane);
        __syncwarp();
        mm_static<32, 16, 32, 2, 16, 32 * 16, 32, 0, 16, 32 * 16>(H, E, G, lane);
      }
      __syncwarp();
      if constexpr (!MMA) {
      // out[l] = sum_(s,B,B') Wc[(s,B,B')][l] H[s][B][B'].  Wc is [(s,B)][B'][l] = 64 rows of 1 KB:
      // lane = (B' mod 8, label quad), so one warp load is 512 contiguous bytes (half a row);
      // 32 loads in flight per lane hide the L2 latency (Wc, 64 KB, does not fit the L1 left
      // beside 213 KB of shared memory).  Issuing the first loads ahead of the H contraction
      // was tried and did not pay (the next image's 22 prefetch registers are live as well).
      const int lq = lane & 3, bpl = lane >> 2;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      const float4* wl = reinterpret_cast<const float4*>(wc) + lane;  // (bp = bpl, lq): float4 index bp * 4 + lq
      const float* hl = H + bpl;
#pragma unroll 1
      for (int r0 = 0; r0 < 64; r0 += 16) {
        float4 wv[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) wv[j] = __ldg(wl + (size_t)(r0 + (j >> 1)) * 64 + (j & 1) * 32);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float h = hl[(r0 + (j >> 1)) * 16 + (j & 1) * 8];
          acc[0] = fmaf(wv[j].x, h, acc[0]);
          acc[1] = fmaf(wv[j].y, h, acc[1]);
          acc[2] = fmaf(wv[j].z, h, acc[2]);
          acc[3] = fmaf(wv[j].w, h, acc[3]);
        }
      }
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] += __shfl_xor_sync(FULL, acc[j], o);
      }
      float* ov = reinterpret_cast<float*>(P.overlaps) + (size_t)img * 16;
      float best = -1.f;
      int besti = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float v = fabsf(acc[j]);
        if (lane < 4) {
          ov[4 * lq + j] = P.signed_out ? acc[j] : v;
          if (v > best) { best = v; besti = 4 * lq + j; }
        }
      }
      if (P.argmax) {
        for (int o = 1; o < 4; o <<= 1) {
          const float ob = __shfl_xor_sync(FULL, best, o);
          const int oi = __shfl_xor_sync(FULL, besti, o);
          if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
        }
        if (lane == 0) P.argmax[img] = besti;
      }
      }  // !MMA
    } else {
    // ---- left sweep.  State Et[B][a] (ld r4(a)); the last step leaves EL[a][B] ----
    float* E = Eb;
    if (lane == 0) E[0] = 1.f;
    __syncwarp();
    for (int n = 0; n < c; ++n) {
      const int al = P.a[n], ar = P.a[n + 1], Bl = P.B[n], Br = P.B[n + 1];
      const int ldE = r4(al), ldW = r4(Br), ldA = r4(ar), ldT = r4(Br);
      const float* W = sW + F.w_off[n];
      const float* A = As + F.a_off[n];
      // T[(s,a)][B'] = sum_B Et[B][a] W[s][B][B']
      mm_slices(Tb, ldT, al * ldT, E, ldE, 0, W, ldW, Bl * ldW, al, Br, Bl, 2, lane);
      __syncwarp();
      if (n + 1 < c) {
        // Et'[B'][a'] = sum_(s,a) T[(s,a)][B'] A[(s,a)][a']
        mm_slices(E, ldA, 0, Tb, ldT, 0, A, ldA, 0, Br, ar, 2 * al, 1, lane);
      } else {
        // EL[a'][B'] = sum_(s,a) A[(s,a)][a'] T[(s,a)][B']
        mm_slices(E, ldT, 0, A, ldA, 0, Tb, ldT, 0, ar, Br, 2 * al, 1, lane);
      }
      __syncwarp();
    }
    // c == 0: EL = [[1]] already (ld irrelevant)

    // ---- right sweep.  State R[a'][B'] (ld r4(B')) ----
    float* R = Rb0;
    float* R2 = Rb1;
    if (lane == 0) R[0] = 1.f;
    __syncwarp();
    for (int n = L - 1; n > c; --n) {
      const int al = P.a[n], ar = P.a[n + 1], Bl = P.B[n], Br = P.B[n + 1];
      const int ldR = r4(Br), ldAt = r4(al), ldWt = r4(Bl), ldTp = r4(al);
      const float* Wt = sW + F.w_off[n];  // [s][B'][B]
      const float* At = As + F.a_off[n];  // [s][a'][a]
      // T'[s][B'][a] = sum_a' R[a'][B'] At[s][a'][a]
      mm_slices(Tb, ldTp, Br * ldTp, R, ldR, 0, At, ldAt, ar * ldAt, Br, al, ar, 2, lane);
      __syncwarp();
      // R'[a][B] = sum_(s,B') T'[(s,B')][a] Wt[(s,B')][B]
      mm_kmajor(R2, ldWt, Tb, ldTp, Wt, ldWt, al, Bl, 2 * Br, lane);
      __syncwarp();
      float* t = R; R = R2; R2 = t;
    }

    // ---- hairy site ----
    {
      const int al = P.a[c], ar = P.a[c + 1], Bl = P.B[c], Br = P.B[c + 1];
      const int ldAt = r4(al), ldR = r4(Br), ldG = r4(Br), ldEL = r4(Bl), ldH = r4(Br);
      const float* At = As + F.a_off[c];  // [s][a'][a]
      float* G = Tb;                      // [s][a][B']  (2 * al * ldG <= 1024)
      float* H = Tb + 1024;               // [s][B][B']  (2 * Bl * ldH <= 1024)
      // G[s][a][B'] = sum_a' At[s][a'][a] ER[a'][B']
      mm_slices(G, ldG, al * ldG, At, ldAt, ar * ldAt, R, ldR, 0, al, Br, ar, 2, lane);
      __syncwarp();
      // H[s][B][B'] = sum_a EL[a][B] G[s][a][B']
      mm_slices(H, ldH, Bl * ldH, E, ldEL, 0, G, ldG, al * ldG, Bl, Br, al, 2, lane);
      __syncwarp();
      // out[l] = sum_(s,B,B') Wc[(s,B,B')][l] H[s][B][B']; lane = (label quad, part)
      const int lq_n = F.Spad >> 2;  // label quads (power of two <= 8)
      const int lq = lane & (lq_n - 1), part = lane / lq_n, parts = 32 / lq_n;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      const int rows = 2 * Bl;
      for (int r = part; r < rows; r += parts) {
        const float* hrow = H + r * ldH;
        const float4* wrow = reinterpret_cast<const float4*>(wc + ((int64_t)r * ldH) * F.Spad) + lq;
#pragma unroll 16
        for (int bp = 0; bp < Br; ++bp) {
          const float h = hrow[bp];
          const float4 w = __ldg(wrow + bp * lq_n);
          acc[0] = fmaf(w.x, h, acc[0]);
          acc[1] = fmaf(w.y, h, acc[1]);
          acc[2] = fmaf(w.z, h, acc[2]);
          acc[3] = fmaf(w.w, h, acc[3]);
        }
      }
      for (int o = lq_n; o < 32; o <<= 1) {
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] += __shfl_xor_sync(FULL, acc[j], o);
      }
      // lanes 0 .. lq_n-1 hold labels 4 lq .. 4 lq + 3
      float* ov = reinterpret_cast<float*>(P.overlaps) + (size_t)img * S;
      float best = -1.f;
      int besti = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int l = 4 * lq + j;
        const float v = fabsf(acc[j]);
        if (lane < lq_n && l < S) {
          ov[l] = P.signed_out ? acc[j] : v;
          if (v > best) { best = v; besti = l; }
        }
      }
      if (P.argmax) {
        // first index of the maximum over lanes 0 .. lq_n-1
        for (int o = 1; o < lq_n; o <<= 1) {
          const float ob = __shfl_xor_sync(FULL, best, o);
          const int oi = __shfl_xor_sync(FULL, besti, o);
          if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
        }
        if (lane == 0) P.argmax[img] = besti;
      }
    }
    }  // !NAT
    __syncwarp();
    }  // valid
    if constexpr (MMA) {
      // ---- label sums of the CTA's 8 images as ONE tensor-core product:
      //      out[l][i] = sum_r Wc[r][l] H_i[r],  r = (s,B,B') < 1024: M = 16 labels, N = 8 images,
      //      warp w takes r in [128 w, 128 w + 128).  Wc (64 KB) is read from L2 once per 8 images
      //      instead of once per image (it was 6 TB/s of L2 traffic and a quarter of the samples).
      __syncthreads();  // every warp's H is in its scratch
      {
        const int g = lane >> 2, t = lane & 3;
        const float* hb = sW + ((F.small_w_total + 31) & ~31) + (size_t)g * PER_WARP + MAXA + kOvE + 1024 +
                          128 * warp;            // H of image g (warp g), this warp's rows
        const float* wcw = wc + (size_t)(128 * warp) * 16;
        float acc4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ks = 0; ks < 16; ++ks) {
          const int k0 = 8 * ks + t, k1 = k0 + 4;
          uint32_t ah[4], al[4], bh[2], bl[2];
          split_tf32(__ldg(wcw + k0 * 16 + g), ah[0], al[0]);
          split_tf32(__ldg(wcw + k0 * 16 + g + 8), ah[1], al[1]);
          split_tf32(__ldg(wcw + k1 * 16 + g), ah[2], al[2]);
          split_tf32(__ldg(wcw + k1 * 16 + g + 8), ah[3], al[3]);
          split_tf32(hb[k0], bh[0], bl[0]);
          split_tf32(hb[k1], bh[1], bl[1]);
          mma_tf32(acc4, al, bh);
          mma_tf32(acc4, ah, bl);
          mma_tf32(acc4, ah, bh);
        }
        // partial sums [label][image] into this warp's (free) G scratch
        float* red = Tb;
        *reinterpret_cast<float2*>(red + g * 8 + 2 * t) = make_float2(acc4[0], acc4[1]);
        *reinterpret_cast<float2*>(red + (g + 8) * 8 + 2 * t) = make_float2(acc4[2], acc4[3]);
      }
      __syncthreads();
      {
        // warp = image: lane l < 16 adds the 8 partial sums of label l
        float v = 0.f;
        if (lane < 16) {
          const float* rb = sW + ((F.small_w_total + 31) & ~31) + MAXA + kOvE + lane * 8 + warp;
#pragma unroll
          for (int w2 = 0; w2 < 8; ++w2) v += rb[(size_t)w2 * PER_WARP];
        }
        if (valid && lane < 16) reinterpret_cast<float*>(P.overlaps)[(size_t)img * 16 + lane] = P.signed_out ? v : fabsf(v);
        v = fabsf(v);
        if (P.argmax) {
          float best = lane < 16 ? v : -1.f;
          int besti = lane;
#pragma unroll
          for (int o = 1; o < 16; o <<= 1) {
            const float ob = __shfl_xor_sync(FULL, best, o);
            const int oi = __shfl_xor_sync(FULL, besti, o);
            if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
          }
          if (valid && lane == 0) P.argmax[img] = besti;
        }
      }
      __syncthreads();  // the partial sums are read before anyone's next G overwrites them
    }
  }
  if constexpr (TC5) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTc5Cols));
  }
}

inline int& overlap_static_shapes() {
  static int v = 1;
  return v;
}
inline int& overlap_mma() {
  static int v = 1;
  return v;
}
inline int& overlap_pad() {
  static int v = 1;
  return v;
}
// diagnostics: upper bound on the warps per CTA (how does the kernel scale with resident warps?)
inline int& overlap_max_warps() {
  static int v = 8;
  return v;
}
// MODE 6: site 4 of the natural 32/32 profile on tcgen05 (measured against MODE 2 in profiles/; off by default)
inline int& overlap_tc5() {
  static int v = 0;
  return v;
}

// Host-side plan; returns false when the shapes do not fit the fast kernel.
inline bool overlap_fast_plan(const OverlapParams& P, OverlapFast& F, bool allow_tc5 = true) {
  if (P.L > OC_MAX_SITES || P.amax > 32 || P.Bmax > 32 || P.S > 32) return false;
  int64_t w = 0, a = 0;
  for (int n = 0; n < P.L; ++n) {
    const int al = P.a[n], ar = P.a[n + 1], Bl = P.B[n], Br = P.B[n + 1];
    F.a_off[n] = a;
    a += (n < P.c) ? 2LL * al * r4(ar) : 2LL * ar * r4(al);
    a = (a + 3) & ~3LL;
    if (n == P.c) {
      F.w_off[n] = -1;
      // G and H live in the two halves of the scratch buffer
      if (2 * al * r4(Br) > 1024 || 2 * Bl * r4(Br) > 1024) return false;
    } else {
      F.w_off[n] = w;
      w += (n < P.c) ? 2LL * Bl * r4(Br) : 2LL * Br * r4(Bl);
      w = (w + 3) & ~3LL;
      // scratch T: left 2*al*r4(Br), right 2*Br*r4(al); env: E r4(al)*... <= 1024; R <= 256
      if (n < P.c && (2 * al * r4(Br) > kOvT || Br * r4(ar) > kOvE || ar * r4(Br) > kOvE)) return false;
      if (n > P.c && (2 * Br * r4(al) > kOvT || al * r4(Bl) > kOvR || ar * r4(Br) > kOvR)) return false;
    }
  }
  F.w4u_off = -1;
  if (allow_tc5 && overlap_tc5() && overlap_static_shapes() && overlap_mma() && overlap_is_natural(P) && P.S == 16) {
    w = (w + 31) & ~31LL;  // 128-byte aligned operand tiles
    F.w4u_off = (int)w;
    w += 2 * kTc5TileFloats;
  }
  if (w > kOvMaxSmallW || a > kOvMaxA || a > 65535 || P.mps_stride > 16384) return false;
  F.small_w_total = (int)w;
  F.a_total = (int)a;
  int sp = 4;
  while (sp < P.S) sp <<= 1;
  F.Spad = sp;
  F.wc_elems = 2LL * P.B[P.c] * r4(P.B[P.c + 1]) * sp;
  return true;
}

inline bool overlap_fast_supported(const OverlapParams& P) {
  OverlapFast F;
  return overlap_fast_plan(P, F);
}

// scratch_dev: at least (small_w_total + wc_elems) floats, 16-byte aligned.
inline size_t overlap_fast_scratch_floats(const OverlapParams& P) {
  OverlapFast F;
  if (!overlap_fast_plan(P, F)) return 0;
  return (size_t)((F.small_w_total + 31) & ~31) + (size_t)F.wc_elems;
}

// do_prep = false: scratch_dev already holds the re-laid classifier of an earlier call with the same
// classifier contents, shapes and options (the caller's cache, see oc_classifier_cache)
inline int overlap_fast_launch(const OverlapParams& P, float* scratch_dev, int sms, cudaStream_t st,
                               bool do_prep = true) {
  OverlapFast F;
  if (!overlap_fast_plan(P, F)) return -1;
  float* small_w = scratch_dev;
  float* wc = scratch_dev + ((F.small_w_total + 31) & ~31);
  int warps = overlap_max_warps() < 8 && overlap_max_warps() > 0 ? overlap_max_warps() : 8;
  const int PER_WARP = F.w4u_off >= 0 ? kTc5PerWarp : kOvPerWarp;
  const size_t wbytes = (size_t)((F.small_w_total + 31) & ~31) * 4;
  const size_t tabbytes = (((size_t)P.mps_stride * 2 + 15) & ~(size_t)15) + 16;
  while (warps > 1 && wbytes + tabbytes + (size_t)warps * PER_WARP * 4 > 222 * 1024) --warps;
  const bool nat = overlap_static_shapes() && overlap_is_natural(P) && F.Spad == 16;
  if (P.mirrored && overlap_is_natural(P)) return -1;  // (never asked for: an untruncated encoding has no sweep)
  // D_encode < 32 against the natural-32 classifier: the image sites are staged zero-padded to the
  // natural profile, so the shared-memory offsets are those of the natural plan
  const bool pad = overlap_static_shapes() && overlap_mma() && overlap_pad() && warps == 8 && F.Spad == 16 &&
                   overlap_is_padded_natural(P);
  const bool last = overlap_static_shapes() && overlap_pad() && F.Spad == 16 && overlap_is_last_natural(P);
  if (pad || last) {
    OverlapParams Pn = P;
    for (int n = 0; n <= 10; ++n) Pn.a[n] = n <= 5 ? (1 << n) : (1 << (10 - n));
    Pn.amax = 32;
    if (!overlap_fast_plan(Pn, F, false)) return -1;  // same small-W block as the first plan (no tcgen05 tile)
  }
  const bool mma = (nat || pad) && overlap_mma() && warps == 8;
  const bool lmma = last && overlap_mma();
  if (do_prep)
    overlap_prep_kernel<float><<<1, 1024, 0, st>>>(reinterpret_cast<const float*>(P.clf), P, F, small_w, wc,
                                                   lmma ? 0x60u : (mma ? 0x10u : 0u));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  const size_t smem = wbytes + tabbytes + (size_t)warps * PER_WARP * 4;
  static PerDeviceFlag attr_flag;
  bool& attr_set = attr_flag.cur();
  if (!attr_set) {
    e = cudaFuncSetAttribute(overlap_fast_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(overlap_fast_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(overlap_fast_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(overlap_fast_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(overlap_fast_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(overlap_fast_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(overlap_fast_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int64_t need = (P.N + warps - 1) / warps;
  const int grid = (int)(need < sms ? need : sms);
  if (lmma)
    overlap_fast_kernel<5><<<grid, warps * 32, smem, st>>>(P, F, small_w, wc);
  else if (last)
    overlap_fast_kernel<4><<<grid, warps * 32, smem, st>>>(P, F, small_w, wc);
  else if (mma && pad)
    overlap_fast_kernel<3><<<grid, warps * 32, smem, st>>>(P, F, small_w, wc);
  else if (mma && F.w4u_off >= 0 && warps == 8)
    overlap_fast_kernel<6><<<grid, warps * 32, smem, st>>>(P, F, small_w, wc);
  else if (mma)
    overlap_fast_kernel<2><<<grid, warps * 32, smem, st>>>(P, F, small_w, wc);
  else if (nat)
    overlap_fast_kernel<1><<<grid, warps * 32, smem, st>>>(P, F, small_w, wc);
  else
    overlap_fast_kernel<0><<<grid, warps * 32, smem, st>>>(P, F, small_w, wc);
  e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

}  // namespace oc
