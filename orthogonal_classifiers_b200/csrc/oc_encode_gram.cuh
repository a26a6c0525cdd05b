// Gram-matrix front end of the half-warp encoder: steps 0..3 of the chain for
// L = 10, D >= 8 (no truncation before bond 4; for 8 <= D < 16 step 3 keeps the D
// largest singular vectors), replacing the head (steps 0-2) and step-3 Jacobi
// kernels of oc_encode_hw.cuh.
//
// With X_n the (2^(n+1) x 2^(9-n)) row unfolding of the padded image
// (tools.py:218-220), step n of the chain is the SVD of
//   M_n[(s,l), c] = sum_i U^(n-1)[i][l] X_n[2i+s][c],
// U^(n-1) = left singular vectors of X_{n-1} (an isometry onto the row space,
// so nothing is lost).  Hence the left singular vectors of M_n are
//   A_n[(s,l), r] = sum_i U^(n-1)[i][l] U^(n)[2i+s][r],
// the singular values are those of X_n, and U^(n) are the eigenvectors of the
// SMALL Gram matrix G_n = X_n X_n^T (2^(n+1) square):
//   G_3 (16 x 16) is accumulated from the pixels in fp64 (products of fp32
//   pixels are exact in fp64), G_2, G_1, G_0 are partial traces of it;
//   G_n = L L^T by diagonally pivoted Cholesky in fp64 (columns in pivot
//   order, rows in place); one-sided Jacobi on the COLUMNS of L in fp32
//   (Veselic-Hari: converged columns are U^(n) Sigma^(n) with relative
//   accuracy, and the pivoting preconditions the iteration: 3.8 sweeps at
//   n = 3 against 7 on the 16 x 64 rows).
// The Jacobi vectors are 16 / 8 / 4 / 2 long instead of 64 / 128 / 256 / 512, and
// the four eigenproblems of an image run side by side in the 16 lanes of its
// half-warp (lanes 0-7 | 8-11 | 12-13 | 14).  Psi_4 = U^(3)^T X_3 goes to the
// global workspace for the step-4 kernel, as before.
#pragma once
#include <algorithm>
#include <type_traits>
#include "oc_encode_hw.cuh"

namespace oc {
namespace gram {

using hw::dot16;
using hw::kHwWarps;
using hw::make_out;
using hw::Out;
using hw::rotate16;

constexpr int kGStride = 256;      // doubles per image in the Gram workspace (G_3, full symmetric)
constexpr double kCholFloor = 1e-13;  // pivots below this fraction of trace(G) end the factorisation

// ---- pixels: four consecutive amplitudes i .. i+3 of the tail-padded image ----
struct PixRow {
  const void* p;
  int dtype, n_pix;
  bool vec, vec8;
};
__device__ __forceinline__ PixRow pix_row(const EncodeParams& P, int64_t img) {
  PixRow r;
  const char* src = reinterpret_cast<const char*>(P.images);
  const int esz = P.in_dtype == 0 ? 4 : (P.in_dtype == 1 ? 8 : 1);
  r.p = src + (size_t)img * P.ld * esz;
  r.dtype = P.in_dtype;
  r.n_pix = P.n_pix;
  r.vec = P.in_dtype == 0 && ((P.ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  r.vec8 = P.in_dtype == 2 && ((P.ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 3) == 0);
  return r;
}
__device__ __forceinline__ float4 pix4(const PixRow& r, int i) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (r.vec && i + 3 < r.n_pix) {
    v = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(r.p) + i));
  } else if (r.vec8 && i + 3 < r.n_pix) {
    const uchar4 u = __ldg(reinterpret_cast<const uchar4*>(reinterpret_cast<const uint8_t*>(r.p) + i));
    v = make_float4(float(u.x) / 255.f, float(u.y) / 255.f, float(u.z) / 255.f, float(u.w) / 255.f);
  } else {
    if (i < r.n_pix) v.x = load_pixel<float>(r.p, r.dtype, i);
    if (i + 1 < r.n_pix) v.y = load_pixel<float>(r.p, r.dtype, i + 1);
    if (i + 2 < r.n_pix) v.z = load_pixel<float>(r.p, r.dtype, i + 2);
    if (i + 3 < r.n_pix) v.w = load_pixel<float>(r.p, r.dtype, i + 3);
  }
  return v;
}

// byte J of w as the amplitude k / 255.f, the same bits as float(k) / 255.f (checked for all 256 k) without
// I2F, division or a table in shared memory (the table's loads were 13 % of the head kernel's samples):
// 0x4B000000 | k is the float 2^23 + k; q = k * (1/255) is corrected once with the exact remainder.
template <int J>
__device__ __forceinline__ float byte_amp(unsigned w) {
  const float f = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7440u | J)) - 8388608.f;
  const float r = 1.f / 255.f;
  const float q = f * r;
  return fmaf(fmaf(-q, 255.f, f), r, q);
}
// MODE 0: fp32 rows, 16-byte aligned quads, n_pix % 4 == 0 -> one predicated LDG.128, no branches
// (so that the loads of a whole image can be in flight together); MODE 2: the same for u8;
// MODE -1: anything else (per-pixel loads).
template <int MODE>
__device__ __forceinline__ float4 pix4m(const PixRow& r, int i) {
  if constexpr (MODE == 0) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < r.n_pix) v = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(r.p) + i));
    return v;
  } else if constexpr (MODE == 2) {
    unsigned w = 0u;
    if (i < r.n_pix) w = __ldg(reinterpret_cast<const unsigned*>(reinterpret_cast<const uint8_t*>(r.p) + i));
    return make_float4(byte_amp<0>(w), byte_amp<1>(w), byte_amp<2>(w), byte_amp<3>(w));
  } else {
    return pix4(r, i);
  }
}
// The same in two halves, so that a prefetched quad can wait in registers in its raw form (one word for
// uint8) and is converted where it is used, not where it is loaded (the conversion would wait for the load).
template <int MODE>
struct RawQuad {
  using type = float4;
};
template <>
struct RawQuad<2> {
  using type = unsigned;
};
template <int MODE>
__device__ __forceinline__ typename RawQuad<MODE>::type pix4raw(const PixRow& r, int i, bool on) {
  if constexpr (MODE == 2) {
    unsigned w = 0u;
    if (on && i < r.n_pix) w = __ldg(reinterpret_cast<const unsigned*>(reinterpret_cast<const uint8_t*>(r.p) + i));
    return w;
  } else {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (on) v = pix4m<MODE>(r, i);
    return v;
  }
}
template <int MODE>
__device__ __forceinline__ float4 pix4cvt(typename RawQuad<MODE>::type raw) {
  if constexpr (MODE == 2) return make_float4(byte_amp<0>(raw), byte_amp<1>(raw), byte_amp<2>(raw), byte_amp<3>(raw));
  else return raw;
}
inline int pix_mode(const EncodeParams& P) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(P.images);
  if ((P.ld & 3) || (P.n_pix & 3)) return -1;
  if (P.in_dtype == 0 && (a & 15) == 0) return 0;
  if (P.in_dtype == 2 && (a & 3) == 0) return 2;
  return -1;
}

// ---- K1a: G_3 = X_3 X_3^T in fp64, one image per half-warp ----
// X_3 sits in shared memory k-major, Xt[k][row], as fp64.  Lane = (ia, kq): the
// 4-row block ia of G against the blocks ia, ia+1, ia+2 (mod 4) - every block
// pair once, the (ia, ia+2) ones twice - over the k-quarter kq; 48 DFMA per six
// 128-bit shared loads.  The four k-quarters are then summed with shuffles.
constexpr int kXtRow = 18;                    // doubles per k (16 rows + pad: 2-way conflicts at most when staging)
constexpr int kXtQuarter = 16 * kXtRow + 2;  // doubles per k-quarter: the +16 B skew keeps the two
                                             // k-quarters of a quarter-warp on different banks
constexpr int kXtDoubles = 4 * kXtQuarter;

__device__ __forceinline__ double shfl_xor_d(double v, int o) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_xor_sync(FULL, lo, o);
  hi = __shfl_xor_sync(FULL, hi, o);
  return __hiloint2double(hi, lo);
}

template <int MODE>
__global__ void __launch_bounds__(kHwWarps * 32, 3) gram_kernel(const __grid_constant__ EncodeParams P,
                                                                double* __restrict__ g_out) {
  extern __shared__ __align__(16) double smem_d[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int l16 = lane & 15, h = lane >> 4;
  const int64_t img = ((int64_t)blockIdx.x * kHwWarps + warp) * 2 + h;
  const bool valid = img < P.N;
  double* Xt = smem_d + (size_t)(warp * 2 + h) * kXtDoubles;
  // stage: coalesced loads (all 16 in flight), then fp64 stores Xt[k][row]
  {
    PixRow src = pix_row(P, valid ? img : 0);
    float4 v[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      v[r] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid) v[r] = pix4m<MODE>(src, 64 * r + 4 * l16);
    }
    // keep all 16 loads ahead of the first conversion (the compiler otherwise converts, i.e. waits,
    // right behind every load)
#pragma unroll
    for (int r = 0; r < 16; ++r) asm volatile("" : "+f"(v[r].x), "+f"(v[r].y), "+f"(v[r].z), "+f"(v[r].w));
    const int k = 4 * l16;
    double* d = Xt + (k >> 4) * kXtQuarter + (k & 15) * kXtRow;
#pragma unroll
    for (int r = 0; r < 16; r += 2) {
      *reinterpret_cast<double2*>(d + r) = make_double2((double)v[r].x, (double)v[r + 1].x);
      *reinterpret_cast<double2*>(d + kXtRow + r) = make_double2((double)v[r].y, (double)v[r + 1].y);
      *reinterpret_cast<double2*>(d + 2 * kXtRow + r) = make_double2((double)v[r].z, (double)v[r + 1].z);
      *reinterpret_cast<double2*>(d + 3 * kXtRow + r) = make_double2((double)v[r].w, (double)v[r + 1].w);
    }
  }
  __syncwarp();
  const int ia = l16 & 3, kq = l16 >> 2;
  const int ib1 = (ia + 1) & 3, ib2 = (ia + 2) & 3;
  double acc0[4][4], acc1[4][4], acc2[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc0[a][b] = acc1[a][b] = acc2[a][b] = 0.0;
  const double* base = Xt + kq * kXtQuarter;
#pragma unroll 2
  for (int kk = 0; kk < 16; ++kk) {
    const double* row = base + kk * kXtRow;
    const double2 a01 = *reinterpret_cast<const double2*>(row + 4 * ia);
    const double2 a23 = *reinterpret_cast<const double2*>(row + 4 * ia + 2);
    const double2 b01 = *reinterpret_cast<const double2*>(row + 4 * ib1);
    const double2 b23 = *reinterpret_cast<const double2*>(row + 4 * ib1 + 2);
    const double2 c01 = *reinterpret_cast<const double2*>(row + 4 * ib2);
    const double2 c23 = *reinterpret_cast<const double2*>(row + 4 * ib2 + 2);
    const double av[4] = {a01.x, a01.y, a23.x, a23.y};
    const double bv[4] = {b01.x, b01.y, b23.x, b23.y};
    const double cv[4] = {c01.x, c01.y, c23.x, c23.y};
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        if (b <= a) acc0[a][b] = fma(av[a], av[b], acc0[a][b]);
        acc1[a][b] = fma(av[a], bv[b], acc1[a][b]);
        acc2[a][b] = fma(av[a], cv[b], acc2[a][b]);
      }
  }
  // sum over the four k-quarters (lanes l16 ^ 4, l16 ^ 8).  Diagonal block: all-reduce of its
  // 10 entries, stored by the kq = 0 lanes.  Off-diagonal blocks: reduce-scatter - after the
  // two exchanges the lane holds ONE fully summed row (a = 2 (kq & 1) + (kq >> 1)) of each
  // block, so every lane stores 2 x 4 entries and their mirrors, no divergence.
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b <= a; ++b) {
      acc0[a][b] += shfl_xor_d(acc0[a][b], 4);
      acc0[a][b] += shfl_xor_d(acc0[a][b], 8);
    }
  const bool b2 = kq & 1, b3 = kq >> 1;
  double f1[4], f2[4];
  {
    double r1[2][4], r2[2][4];
#pragma unroll
    for (int ap = 0; ap < 2; ++ap)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        r1[ap][b] = (b2 ? acc1[ap + 2][b] : acc1[ap][b]) + shfl_xor_d(b2 ? acc1[ap][b] : acc1[ap + 2][b], 4);
        r2[ap][b] = (b2 ? acc2[ap + 2][b] : acc2[ap][b]) + shfl_xor_d(b2 ? acc2[ap][b] : acc2[ap + 2][b], 4);
      }
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      f1[b] = (b3 ? r1[1][b] : r1[0][b]) + shfl_xor_d(b3 ? r1[0][b] : r1[1][b], 8);
      f2[b] = (b3 ? r2[1][b] : r2[0][b]) + shfl_xor_d(b3 ? r2[0][b] : r2[1][b], 8);
    }
  }
  if (valid) {
    double* g = g_out + (size_t)img * kGStride;
    if (kq == 0) {
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b <= a; ++b) {
          g[(4 * ia + a) * 16 + 4 * ia + b] = acc0[a][b];
          g[(4 * ia + b) * 16 + 4 * ia + a] = acc0[a][b];
        }
    }
    const int ra = 4 * ia + 2 * (kq & 1) + (kq >> 1);
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      g[ra * 16 + 4 * ib1 + b] = f1[b];
      g[(4 * ib1 + b) * 16 + ra] = f1[b];
      g[ra * 16 + 4 * ib2 + b] = f2[b];
      g[(4 * ib2 + b) * 16 + ra] = f2[b];
    }
  }
}

// ---- K1a for uint8 pixels: G_3 = X_3 X_3^T as an EXACT integer Gram matrix on the tensor cores ----
// Pixels k = 0..255 (the amplitude is k / 255): a product is < 2^16 and a sum of 64 of them < 2^22, so
// G_3 * 255^2 is an integer that mma.sync.m16n8k32.u8.u8.s32 (SASS IMMA.16832) accumulates exactly - no
// fp64 staging, no DFMA: the kernel is a stream over the pixels.  One image per warp and pass: lane
// (g, t) = (lane / 4, lane % 4) loads the 16 bytes 16 t .. 16 t + 15 of rows g and g + 8 of X_3 (16 rows
// of 64 pixels), the whole 1 KB image in two 128-bit loads per lane.  G = sum_k X[m][k] X[n][k] does not
// care in which order k is summed, so the lane's four words serve as the k-slots 4 t .. 4 t + 3 of four
// k-blocks (two per MMA), and since B[k][n] = X[n][k] the B fragments ARE the A registers of rows g
// (columns 0-7 of G) and g + 8 (columns 8-15).  Four MMAs per image; the accumulators (rows g, g + 8,
// columns 2 t, 2 t + 1 of both column blocks) go out as fp64 / 255^2, 16 bytes per store.
// Needs 16-byte aligned rows (ld and n_pix multiples of 16: 784 and 1024 are); otherwise gram_kernel<2>.
__device__ __forceinline__ void imma_u8(int (&c)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0,
                                        unsigned b1) {
  asm("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
constexpr int kGramU8Unroll = 4;  // images in flight per warp
__global__ void __launch_bounds__(256) gram_u8_kernel(const __grid_constant__ EncodeParams P, double* __restrict__ g_out) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const uint8_t* images = reinterpret_cast<const uint8_t*>(P.images);
  const int o0 = 64 * g + 16 * t, o1 = o0 + 512;
  const bool in0 = o0 < P.n_pix, in1 = o1 < P.n_pix;  // n_pix % 16 == 0: a 16-byte piece is inside or padding
  const double sc = 1.0 / 65025.0;
  for (int64_t i0 = warp * kGramU8Unroll; i0 < P.N; i0 += nwarps * kGramU8Unroll) {
    uint4 r0[kGramU8Unroll], r1[kGramU8Unroll];
#pragma unroll
    for (int u = 0; u < kGramU8Unroll; ++u) {
      r0[u] = r1[u] = make_uint4(0u, 0u, 0u, 0u);
      if (i0 + u < P.N) {
        const uint8_t* src = images + (size_t)(i0 + u) * P.ld;
        if (in0) r0[u] = __ldcs(reinterpret_cast<const uint4*>(src + o0));
        if (in1) r1[u] = __ldcs(reinterpret_cast<const uint4*>(src + o1));
      }
    }
#pragma unroll
    for (int u = 0; u < kGramU8Unroll; ++u) {
      if (i0 + u >= P.N) break;
      int c0[4] = {0, 0, 0, 0}, c1[4] = {0, 0, 0, 0};
      imma_u8(c0, r0[u].x, r1[u].x, r0[u].y, r1[u].y, r0[u].x, r0[u].y);
      imma_u8(c1, r0[u].x, r1[u].x, r0[u].y, r1[u].y, r1[u].x, r1[u].y);
      imma_u8(c0, r0[u].z, r1[u].z, r0[u].w, r1[u].w, r0[u].z, r0[u].w);
      imma_u8(c1, r0[u].z, r1[u].z, r0[u].w, r1[u].w, r1[u].z, r1[u].w);
      double* G = g_out + (size_t)(i0 + u) * kGStride;
      *reinterpret_cast<double2*>(G + g * 16 + 2 * t) = make_double2(c0[0] * sc, c0[1] * sc);
      *reinterpret_cast<double2*>(G + (g + 8) * 16 + 2 * t) = make_double2(c0[2] * sc, c0[3] * sc);
      *reinterpret_cast<double2*>(G + g * 16 + 8 + 2 * t) = make_double2(c1[0] * sc, c1[1] * sc);
      *reinterpret_cast<double2*>(G + (g + 8) * 16 + 8 + 2 * t) = make_double2(c1[2] * sc, c1[3] * sc);
    }
  }
}
inline int& u8_imma() {
  static int v = 1;
  return v;
}

// ---- shared-memory map of one image in K1b ----
// Gram matrices, row stride p + 1 (lane = row reads column piv without conflicts)
__host__ __device__ constexpr int g_off(int n) { return n == 3 ? 0 : (n == 2 ? 272 : (n == 1 ? 344 : 364)); }
constexpr int kGsDoubles = 372;
// L rows for the pivot-row broadcast, [row][col], stride p
__host__ __device__ constexpr int l_off(int n) { return n == 3 ? 0 : (n == 2 ? 256 : (n == 1 ? 320 : 336)); }
constexpr int kLsDoubles = 340;
// fp32 L columns [k][row], stride p, for the Jacobi load (offsets l_off)
constexpr int kLfFloats = 340;
// After the factorisation the Gram region holds U^(n) as fp32 [rank][row], stride
// p + 1 (offsets g_off), and the L-row region U^(3) transposed, [row][rank].
constexpr int kHeadSmemBytes = kGsDoubles * 8 + kLsDoubles * 8 + kLfFloats * 4;  // 7,056 B: 4 CTAs of 8 images per SM

__device__ __forceinline__ double shfl_d(double v, int src) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_sync(FULL, lo, src);
  hi = __shfl_sync(FULL, hi, src);
  return __hiloint2double(hi, lo);
}

// Diagonally pivoted Cholesky of the p x p matrix g (row stride p + 1) by the p
// lanes [gbase, gbase + p) of the warp, lane - gbase = row.  Lanes of different
// groups (and different p <= KMAX) run side by side; p = 0: idle lane.
// ls: [row][col] fp64 copy (pivot rows are broadcast through it); lf: fp32
// [col][row].  Returns the number of columns (the numerical rank).
template <int KMAX>
__device__ __forceinline__ int pivoted_cholesky(const double* __restrict__ g, double* __restrict__ ls,
                                                float* __restrict__ lf, int p, int gbase, int lane,
                                                double inv_trace) {
  const int row = lane - gbase;
  double lrow[KMAX];
  double d = (p > 0) ? g[row * (p + 1) + row] : 0.0;
  bool done = p == 0, fin = p == 0;
  int rank = 0;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const bool on = k < p && !fin;
    // pivot = largest remaining diagonal (fp32 key, row index in the low bits; ties -> lowest row)
    const float dn = (float)(d * inv_trace);
    unsigned key = 0u;
    if (on && !done && dn > (float)kCholFloor) key = (__float_as_uint(dn) & ~15u) | (unsigned)(15 - row);
    unsigned kmax = key;
#pragma unroll
    for (int o = KMAX / 2; o > 0; o >>= 1) {
      const unsigned other = __shfl_xor_sync(FULL, kmax, o);
      if (o < p) kmax = max(kmax, other);
    }
    if (kmax == 0u) fin = true;
    const int piv = 15 - (int)(kmax & 15u);
    const double dpiv = shfl_d(d, gbase + (piv & (p > 0 ? p - 1 : 0)));
    double v = 0.0;
    if (on && !fin) {
      const double inv = rsqrt(dpiv);
      double dotv = 0.0;
      const double* lp = ls + piv * p;
#pragma unroll
      for (int j = 0; j + 1 < k; j += 2) {
        const double2 b = *reinterpret_cast<const double2*>(lp + j);
        dotv = fma(lrow[j], b.x, dotv);
        dotv = fma(lrow[j + 1], b.y, dotv);
      }
      if (k & 1) dotv = fma(lrow[k - 1], lp[k - 1], dotv);
      v = (g[row * (p + 1) + piv] - dotv) * inv;
      if (row == piv) v = dpiv * inv;
      if (done) v = 0.0;
      ++rank;
    }
    lrow[k] = v;
    if (k < p) {
      ls[row * p + k] = v;
      lf[k * p + row] = (float)v;
    }
    d = fma(-v, v, d);
    if (on && !fin && row == piv) done = true;
    __syncwarp();
  }
  return rank;
}

// One-sided Jacobi (odd-even ordering) on the 2 * gsz vectors of each lane
// group [gbase, gbase + gsz), 16 elements per vector, groups of different size
// side by side.  Same rotation kernel and stopping rule as hw::jacobi16.
// NEJ packed pairs per vector take part (7 when rows 14, 15 of X_3 are padding: n_pix <= 896).
template <int NEJ>
__device__ __forceinline__ int jacobi_groups(u64 (&T)[16], u64 (&S)[16], float& nT, float& nS, float thr2,
                                             int lane, int gsz, int gbase, int mlive, int mlive_max) {
  const int slot = lane - gbase;
  const int src_dn = gbase + ((slot + 1) & (gsz - 1));
  const int src_up = gbase + ((slot - 1) & (gsz - 1));
  const bool idleB = slot >= mlive - 1;
  const unsigned gmask = ((1u << gsz) - 1u) << gbase;
  int sweeps = 0;
  bool done = mlive == 0;
#pragma unroll 1
  for (int sw = 0; sw < kMaxSweeps; ++sw) {
    bool again = false;
    bool lostf = false;
    {
      const float fT = dot16<0, NEJ>(T, T), fS = dot16<0, NEJ>(S, S);
      if (!done) {
        nT = fT;
        nS = fS;
      }
    }
#pragma unroll 1
    for (int rr = 0; rr < mlive_max; ++rr) {
      const bool off = done || rr >= mlive;
      rotate16<0, NEJ>(T, S, nT, nS, thr2, off, lostf, again);
#pragma unroll
      for (int e = 0; e < NEJ; ++e) T[e] = shfl2(T[e], src_dn);
      float nR = __shfl_sync(FULL, nT, src_dn);
      rotate16<0, NEJ, true>(S, T, nS, nR, thr2, off || idleB, lostf, again);
#pragma unroll
      for (int e = 0; e < NEJ; ++e) T[e] = shfl2(T[e], src_up);
      nT = __shfl_sync(FULL, nR, src_up);
      {
#ifdef OC_TRACK_LOST
        const unsigned lostmask = __ballot_sync(FULL, lostf);
#else
        const unsigned lostmask = 0u;
#endif
        if (lostmask) {
          const float fT = dot16<0, NEJ>(T, T), fS = dot16<0, NEJ>(S, S);
          if (lostmask & gmask) {
            nT = fT;
            nS = fS;
          }
          lostf = false;
        }
      }
    }
    if (!done) ++sweeps;
    const unsigned b = __ballot_sync(FULL, again);
    if (!(b & gmask)) done = true;  // also for two vectors: one fp32 rotation leaves cos ~ 1e-4 when graded
    if (__all_sync(FULL, done)) break;
  }
  return sweeps;
}

// ---- K1b: steps 0..3 from G_3; Psi_4 to the workspace ----
template <int MODE, int NEJ>
__global__ void __launch_bounds__(kHwWarps * 32, 4)
gram_head_kernel(const __grid_constant__ EncodeParams P, const double* __restrict__ g_in,
                 float* __restrict__ ws_out, uint8_t* __restrict__ keys) {
  extern __shared__ __align__(16) unsigned char smem_b[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int l16 = lane & 15, hb = lane & 16, h = lane >> 4;
  const int64_t img = ((int64_t)blockIdx.x * kHwWarps + warp) * 2 + h;
  const bool valid = img < P.N;
  const int64_t im = valid ? img : 0;
  unsigned char* mine = smem_b + (size_t)(warp * 2 + h) * kHeadSmemBytes;
  double* Gs = reinterpret_cast<double*>(mine);
  double* Ls = Gs + kGsDoubles;
  float* Lf = reinterpret_cast<float*>(Ls + kLsDoubles);
  float* Uf = reinterpret_cast<float*>(Gs);  // after the factorisation
  float* Ut = reinterpret_cast<float*>(Ls);  // after the factorisation

  // ---- G_3 -> shared (row stride 17), partial traces G_2, G_1, G_0 ----
  {
    const double* g = g_in + (size_t)im * kGStride;
#pragma unroll
    for (int e = l16; e < 256; e += 16) Gs[(e >> 4) * 17 + (e & 15)] = valid ? __ldg(g + e) : 0.0;
  }
  __syncwarp();
  double tr = Gs[l16 * 17 + l16];
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) tr += shfl_d(tr, lane ^ o);
  const double inv_trace = tr > 0.0 ? 1.0 / tr : 0.0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int e = l16 + 16 * q, i = e >> 3, j = e & 7;
    Gs[g_off(2) + i * 9 + j] = Gs[(2 * i) * 17 + 2 * j] + Gs[(2 * i + 1) * 17 + 2 * j + 1];
  }
  __syncwarp();
  {
    const int i = l16 >> 2, j = l16 & 3;
    Gs[g_off(1) + i * 5 + j] =
        Gs[g_off(2) + (2 * i) * 9 + 2 * j] + Gs[g_off(2) + (2 * i + 1) * 9 + 2 * j + 1];
  }
  __syncwarp();
  if (l16 < 4) {
    const int i = l16 >> 1, j = l16 & 1;
    Gs[g_off(0) + i * 3 + j] =
        Gs[g_off(1) + (2 * i) * 5 + 2 * j] + Gs[g_off(1) + (2 * i + 1) * 5 + 2 * j + 1];
  }
  __syncwarp();

  // ---- pivoted Cholesky: problem 3 on all 16 lanes, then 2 | 1 | 0 side by side ----
  const int rank3 = pivoted_cholesky<16>(Gs, Ls, Lf, 16, hb, lane, inv_trace);
  // small problems: lanes 0-7 -> n = 2, 8-11 -> n = 1, 12-13 -> n = 0, 14-15 idle
  const int nS_ = l16 < 8 ? 2 : (l16 < 12 ? 1 : 0);
  const int pS_ = l16 < 8 ? 8 : (l16 < 12 ? 4 : (l16 < 14 ? 2 : 0));
  const int bS_ = hb + (l16 < 8 ? 0 : (l16 < 12 ? 8 : (l16 < 14 ? 12 : l16)));
  const int rankS = pivoted_cholesky<8>(Gs + g_off(nS_), Ls + l_off(nS_), Lf + l_off(nS_), pS_, bS_, lane,
                                        inv_trace);
  const int rank2 = __shfl_sync(FULL, rankS, hb + 0);
  const int rank1 = __shfl_sync(FULL, rankS, hb + 8);
  const int rank0 = __shfl_sync(FULL, rankS, hb + 12);

  // ---- Jacobi on the columns of L_n: lanes 0-7 -> n = 3 (8 slots), 8-11 -> 2, 12-13 -> 1, 14 -> 0 ----
  const int nJ = l16 < 8 ? 3 : (l16 < 12 ? 2 : (l16 < 14 ? 1 : 0));
  const int gsz = l16 < 8 ? 8 : (l16 < 12 ? 4 : (l16 < 14 ? 2 : 1));
  const int gb16 = l16 < 8 ? 0 : (l16 < 12 ? 8 : (l16 < 14 ? 12 : l16));
  const int gbase = hb + gb16;
  const int slot = l16 - gb16;
  const int pJ = 2 * gsz;  // matrix size of this lane's problem
  const int rankJ = l16 == 15 ? 0 : (nJ == 3 ? rank3 : (nJ == 2 ? rank2 : (nJ == 1 ? rank1 : rank0)));
  u64 T[16], S[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) T[e] = S[e] = 0ull;
  {
    const float* lt = Lf + l_off(nJ) + (2 * slot) * pJ;
    const float* lsn = lt + pJ;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (2 * e < pJ && l16 != 15) {
        const float2 a = *reinterpret_cast<const float2*>(lt + 2 * e);
        const float2 b = *reinterpret_cast<const float2*>(lsn + 2 * e);
        T[e] = pack2(a.x, a.y);
        S[e] = pack2(b.x, b.y);
      }
    }
  }
  __syncwarp();
  float nT = dot16<0, 8>(T, T), nS = dot16<0, 8>(S, S);
  float fro = nT + nS;
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) {
    const float v = __shfl_xor_sync(FULL, fro, o);
    if (o < gsz) fro += v;
  }
  const float thr2 = kFloor2 * fro;
  const int mlive = (rankJ + 1) >> 1;
  int mlive_max = mlive;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mlive_max = max(mlive_max, __shfl_xor_sync(FULL, mlive_max, o));
  int sweeps = 0;
  if (mlive_max > 0) {
    sweeps = jacobi_groups<NEJ>(T, S, nT, nS, thr2, lane, gsz, gbase, mlive, mlive_max);
    nT = dot16<0, 8>(T, T);
    nS = dot16<0, 8>(S, S);
  }
  // stable descending rank among the 2 * gsz vectors of the group
  int rT = 0, rS = 0;
  {
    const int pT = 2 * slot, pS = pT + 1;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float v = __shfl_sync(FULL, (j & 1) ? nS : nT, gbase + ((j >> 1) & (gsz - 1)));
      if (j < pJ) {
        rT += (v > nT) || (v == nT && j < pT);
        rS += (v > nS) || (v == nS && j < pS);
      }
    }
  }
  const Out io = make_out(P, im, nJ);
  const int r_rt = io.r_rt;
  const bool aT = nT > thr2 && l16 != 15, aS = nS > thr2 && l16 != 15;
  const bool kT = aT && rT < r_rt, kS = aS && rS < r_rt;
  if (valid && l16 != 15) {
    if (io.sv) {
      if (rT < io.sv_width) io.sv[rT] = aT ? sqrtf(nT) : 0.f;
      if (rS < io.sv_width) io.sv[rS] = aS ? sqrtf(nS) : 0.f;
      for (int k = pJ + slot; k < io.sv_width; k += gsz) io.sv[k] = 0.f;
    }
    if (io.sweeps && slot == 0) *io.sweeps = sweeps;
  }
  int kept_max;  // of the warp's two images: columns >= kept of U^(3) are zero, the rows of Psi_4 behind them too
  {
    // kept vectors of problem 3 = non-zero rows of Psi_4 (pairing key of step 4)
    const unsigned kt = __ballot_sync(FULL, kT && nJ == 3), ks = __ballot_sync(FULL, kS && nJ == 3);
    const int kept = __popc((kt >> hb) & 0xffu) + __popc((ks >> hb) & 0xffu);
    if (keys && valid && l16 == 0) keys[im] = (uint8_t)kept;
    kept_max = max(__popc(kt & 0xffu) + __popc(ks & 0xffu), __popc((kt >> 16) & 0xffu) + __popc((ks >> 16) & 0xffu));
  }
  // ---- U^(n)[row][rank] -> shared: Uf [rank][row] (stride p + 1); U^(3) also transposed ----
  if (l16 != 15) {
    const float iT = kT ? rsqrtf(nT) : 0.f, iS = kS ? rsqrtf(nS) : 0.f;
    float* uT = Uf + g_off(nJ) + rT * (pJ + 1);
    float* uS = Uf + g_off(nJ) + rS * (pJ + 1);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (2 * e < pJ) {
        float t0, t1, s0, s1;
        unpack2(T[e], t0, t1);
        unpack2(S[e], s0, s1);
        t0 *= iT; t1 *= iT; s0 *= iS; s1 *= iS;
        uT[2 * e] = t0; uT[2 * e + 1] = t1;
        uS[2 * e] = s0; uS[2 * e + 1] = s1;
        if (nJ == 3) {
          Ut[(2 * e) * 16 + rT] = t0; Ut[(2 * e + 1) * 16 + rT] = t1;
          Ut[(2 * e) * 16 + rS] = s0; Ut[(2 * e + 1) * 16 + rS] = s1;
        }
      }
    }
  }
  __syncwarp();

  // first rows of X_3 for the Psi_4 product: in flight during the A_n products
  PixRow src = pix_row(P, im);
  const int rows_live = (P.n_pix + 63) >> 6;
  typename RawQuad<MODE>::type xa[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) xa[j] = pix4raw<MODE>(src, 64 * j + 4 * l16, valid && j < rows_live);

  // ---- A_n[(s,l), r] = sum_i U^(n-1)[i][l] U^(n)[2i+s][r]  (D >= 8: bonds 1..3 are the natural 2, 4, 8) ----
  if (valid) {
    float* mps = reinterpret_cast<float*>(P.mps_out) + (size_t)im * P.mps_stride;
    if (l16 < 4) {  // n = 0: A_0[(s,0), r] = U^(0)[s][r]
      const int s = l16 >> 1, r = l16 & 1;
      mps[P.site_off[0] + s * 2 + r] = Uf[g_off(0) + r * 3 + s];
    }
    {  // n = 1: one output per lane, (s, l, r) = (bit 3, bit 2, bits 1-0)
      const int r = l16 & 3, l = (l16 >> 2) & 1, s = l16 >> 3;
      const float* up = Uf + g_off(0) + l * 3;
      const float* un = Uf + g_off(1) + r * 5 + s;
      mps[P.site_off[1] + (s * 2 + l) * 4 + r] = fmaf(up[0], un[0], up[1] * un[2]);
    }
    {  // n = 2: lane = (s, r), four l each
      const int r = l16 & 7, s = l16 >> 3;
      const float* un = Uf + g_off(2) + r * 9 + s;
      const float c0 = un[0], c1 = un[2], c2 = un[4], c3 = un[6];
      float* A = mps + P.site_off[2] + (s * 4) * 8 + r;
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        const float* up = Uf + g_off(1) + l * 5;
        A[l * 8] = fmaf(up[0], c0, fmaf(up[1], c1, fmaf(up[2], c2, up[3] * c3)));
      }
    }
    {  // n = 3: lane = r, all (s, l); bond 4 = min(16, D) columns (8 <= D < 16 truncates here)
      const int r4b = P.bonds[4];
      const float* un = Uf + g_off(3) + l16 * 17;
      float col[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) col[j] = un[j];
      float* A = mps + P.site_off[3] + l16;
#pragma unroll
      for (int l = 0; l < 8; ++l) {
        const float* up = Uf + g_off(2) + l * 9;
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float u = up[i];
          a0 = fmaf(u, col[2 * i], a0);
          a1 = fmaf(u, col[2 * i + 1], a1);
        }
        if (l16 < r4b) {
          A[l * r4b] = a0;
          A[(8 + l) * r4b] = a1;
        }
      }
    }
  }

  // ---- Psi_4[l][c] = sum_i U^(3)[i][l] X_3[i][c]; this lane: c = 4 l16 .. 4 l16 + 3, all l ----
  {
    float acc[16][4];
#pragma unroll
    for (int l = 0; l < 16; ++l)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[l][q] = 0.f;
#pragma unroll 1
    for (int i0 = 0; i0 < rows_live; i0 += 4) {
      typename RawQuad<MODE>::type xb[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        xb[j] = pix4raw<MODE>(src, 64 * (i0 + 4 + j) + 4 * l16, valid && i0 + 4 + j < rows_live);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 xq = pix4cvt<MODE>(xa[j]);
        const float xv[4] = {xq.x, xq.y, xq.z, xq.w};
#pragma unroll
        for (int l4 = 0; l4 < 4; ++l4) {
          if (4 * l4 >= kept_max) continue;  // warp-uniform: zero columns of U^(3) (13 live rows: rank <= 13)
          const float4 u = *reinterpret_cast<const float4*>(Ut + (i0 + j) * 16 + 4 * l4);
          const float uv[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[4 * l4 + a][q] = fmaf(uv[a], xv[q], acc[4 * l4 + a][q]);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) xa[j] = xb[j];
    }
    if (valid) {
      float* out = ws_out + (size_t)im * kWsStride + 4 * l16;
#pragma unroll
      for (int l = 0; l < 16; ++l)
        *reinterpret_cast<float4*>(out + l * 64) = make_float4(acc[l][0], acc[l][1], acc[l][2], acc[l][3]);
    }
  }
}

inline int& enabled() {
  static int v = 1;
  return v;
}

// steps 0..3: images -> A_0..A_3, Psi_4 (ws_psi4), keys (nullable); gws: kGStride doubles per image
inline cudaError_t launch_head(const EncodeParams& P, double* gws, float* ws_psi4, uint8_t* keys,
                               cudaStream_t st) {
  const int per_cta = kHwWarps * 2;
  const int grid = (int)((P.N + per_cta - 1) / per_cta);
  const size_t smem_a = (size_t)per_cta * kXtDoubles * sizeof(double);
  const size_t smem_b = (size_t)per_cta * kHeadSmemBytes;
  const int mode = pix_mode(P);
  auto launch = [&](auto mode_tag) -> cudaError_t {
    constexpr int MODE = decltype(mode_tag)::value;
    static PerDeviceFlag attr_flag;
    bool& attr_set = attr_flag.cur();
    if (!attr_set) {
      cudaError_t e = cudaFuncSetAttribute(gram_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a);
      if (e != cudaSuccess) return e;
      e = cudaFuncSetAttribute(gram_head_kernel<MODE, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b);
      if (e != cudaSuccess) return e;
      e = cudaFuncSetAttribute(gram_head_kernel<MODE, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b);
      if (e != cudaSuccess) return e;
      attr_set = true;
    }
    if (MODE == 2 && u8_imma() && (P.ld & 15) == 0 && (P.n_pix & 15) == 0 &&
        (reinterpret_cast<uintptr_t>(P.images) & 15) == 0) {
      const int64_t want = (P.N + 8 * kGramU8Unroll - 1) / (8 * kGramU8Unroll);
      gram_u8_kernel<<<(int)std::min<int64_t>(want, 148 * 8), 256, 0, st>>>(P, gws);
    } else {
      gram_kernel<MODE><<<grid, kHwWarps * 32, smem_a, st>>>(P, gws);
    }
    // rows 14 and 15 of X_3 (pixels 896..1023) are padding for MNIST-shaped images: the L columns end there
    if (P.n_pix <= 896) gram_head_kernel<MODE, 7><<<grid, kHwWarps * 32, smem_b, st>>>(P, gws, ws_psi4, keys);
    else gram_head_kernel<MODE, 8><<<grid, kHwWarps * 32, smem_b, st>>>(P, gws, ws_psi4, keys);
    return cudaGetLastError();
  };
  if (mode == 0) return launch(std::integral_constant<int, 0>{});
  if (mode == 2) return launch(std::integral_constant<int, 2>{});
  return launch(std::integral_constant<int, -1>{});
}

}  // namespace gram
}  // namespace oc
