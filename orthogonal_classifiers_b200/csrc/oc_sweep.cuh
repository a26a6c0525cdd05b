// OC_TRUNC_SWEEP: the truncation order of xmps' `left_canonicalise(D)` as the reference calls it
// (tools.py:221): exact left-to-right chain, then a RIGHT-TO-LEFT SVD sweep that truncates every
// bond to min(D, rank), then a sweep back to the left-canonical gauge.
//
// A right-to-left truncating sweep over the exact left-canonical MPS sees, at every bond, the true
// Schmidt values of the state already truncated on the bonds to its right: as a map on STATES it is
// the inline truncated SVD chain run from the right end, i.e. the inline chain of the image with
// the bits of its flat pixel index reversed (site n <-> site L-1-n).  So:
//
//   bitrev_kernel      x'[j] = x[bitrev_L(j)]  (tail padding and /255 applied on the way)
//   (inline encoder)   A'_n (2, a_n, a_{n+1}) of x' - any of the encoders of this library
//   sweep_back_kernel  B_m[s][l][r] = A'_{L-1-m}[s][r][l] is a RIGHT-canonical MPS of the truncated
//                      state; one left-to-right pass of QR factorisations (no truncation) brings it
//                      to the left-canonical gauge the reference returns.  The reference does this
//                      pass with SVDs (U kept, S.V carried); QR gives the same state in another gauge
//                      of the bond spaces, and site tensors are gauge-dependent anyway.
//
// QR per site: M[(s,k), r] = sum_l C[k][l] B_m[s][l][r] (C = carry, 1 x 1 at site 0); modified
// Gram-Schmidt on the columns, done TWICE ("twice is enough": the columns of M are graded like the
// Schmidt spectrum, one pass loses orthogonality like eps * sigma_max / sigma_min, the second pass
// works on a matrix with condition ~1); A_m = Q, carry <- R2 R1.  One warp per image, lane = column
// (<= 32), the column in registers, q broadcast through shared memory.  Columns whose residual is
// below floor * |M|_F are dropped (zero column, as the encoders do for rank-deficient images) and
// the live columns are compacted to the front, so zero columns trail as in the other encoders.
// The last carry (1 x 1, = 1 for a unit-norm state) is dropped as the reference drops it.
#pragma once
#include "oc_generic.cuh"

namespace oc {
namespace sweep {

template <typename T>
__global__ void __launch_bounds__(256) bitrev_kernel(const void* __restrict__ images, int in_dtype, int64_t N,
                                                     int n_pix, int64_t ld, int L, T* __restrict__ out) {
  __shared__ T buf[1024 + 32];
  const int cap = 1 << L;
  const int esz = in_dtype == 0 ? 4 : (in_dtype == 1 ? 8 : 1);
  for (int64_t img = blockIdx.x; img < N; img += gridDim.x) {
    const void* row = reinterpret_cast<const char*>(images) + (size_t)img * ld * esz;
    for (int i = threadIdx.x; i < cap; i += blockDim.x)
      buf[i + (i >> 5)] = i < n_pix ? load_pixel<T>(row, in_dtype, i) : T(0);
    __syncthreads();
    T* o = out + (size_t)img * cap;
    for (int j = threadIdx.x; j < cap; j += blockDim.x) {
      const int i = (int)(__brev((unsigned)j) >> (32 - L));
      o[j] = buf[i + (i >> 5)];
    }
    __syncthreads();
  }
}

struct SweepParams {
  const void* tmp_mps;   // inline encoding of the bit-reversed images (fixed natural layout)
  void* mps_out;         // left-canonical result, same layout
  const void* tmp_sv;    // nullable: svals rows of the mirrored chain
  void* sv_out;          // nullable
  const int32_t* tmp_sw; // nullable: Jacobi sweep counts of the mirrored chain
  int32_t* sw_out;       // nullable
  int64_t N;
  int L;
  int sv_width;
  int64_t mps_stride;
  int64_t site_off[OC_MAX_L + 1];
  int bonds[OC_MAX_L + 1];
};

// Per-image shared-memory block for lane groups of G lanes (G = 8, 16, 32 >= the largest bond):
// Bs 2 G (G + 1) (dead after the M build: R1 | R2 reuse it) | pad | C G^2 | q 2 G   (elements of T)
__host__ __device__ constexpr int sw_per_image(int G) { return 3 * G * G + 4 * G; }

template <typename T, int G>
__device__ __forceinline__ T group_sum_g(T v) {
#pragma unroll
  for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// sum_i a[i] * b[i] with four independent accumulators (a serial chain of ROWS dependent FMAs
// would cost ROWS x 4 cycles per step with only a few warps per scheduler to hide it)
// q (ROWS elements, 16-byte aligned shared memory) -> registers with 128-bit loads
template <typename T, int ROWS>
__device__ __forceinline__ void load_q(T (&q)[ROWS], const T* __restrict__ qb) {
  if constexpr (sizeof(T) == 4) {
#pragma unroll
    for (int i = 0; i < ROWS; i += 4) {
      const float4 v = *reinterpret_cast<const float4*>(qb + i);
      q[i] = v.x; q[i + 1] = v.y; q[i + 2] = v.z; q[i + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < ROWS; i += 2) {
      const double2 v = *reinterpret_cast<const double2*>(qb + i);
      q[i] = v.x; q[i + 1] = v.y;
    }
  }
}
template <typename T, int ROWS>
__device__ __forceinline__ void store_q(T* __restrict__ qb, const T (&q)[ROWS]) {
  if constexpr (sizeof(T) == 4) {
#pragma unroll
    for (int i = 0; i < ROWS; i += 4) *reinterpret_cast<float4*>(qb + i) = make_float4(q[i], q[i + 1], q[i + 2], q[i + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < ROWS; i += 2) *reinterpret_cast<double2*>(qb + i) = make_double2(q[i], q[i + 1]);
  }
}

template <typename T, int ROWS, typename A>
__device__ __forceinline__ T dot4(const A& a, const T (&b)[ROWS]) {
  T s0 = T(0), s1 = T(0), s2 = T(0), s3 = T(0);
#pragma unroll
  for (int i = 0; i < ROWS; i += 4) {
    s0 += a[i] * b[i];
    s1 += a[i + 1] * b[i + 1];
    s2 += a[i + 2] * b[i + 2];
    s3 += a[i + 3] * b[i + 3];
  }
  return (s0 + s1) + (s2 + s3);
}

// One site for every image of the warp (32 / G images, G lanes each, lane = column), with
// ROWS = 2 * HR register rows per column (HR >= a_m).  Loop bounds are warp-uniform (the bond
// profile is fixed); what depends on the image (dead columns, rank) is predicated.
template <typename T, int ROWS, int G>
__device__ __forceinline__ void qr_site(const T* __restrict__ Ain, T* __restrict__ Aout, bool valid, int al, int ar,
                                        T* Bs, T* C, T* R1, T* R2, T* qb, int gl, bool last) {
  constexpr int HR = ROWS / 2;
  // ---- stage B_m[s][l][r] = A'[s][r][l] as Bs[(s * al + l) * (G + 1) + r]: lane = l, rows of A'
  //      read contiguously, the odd stride keeps the transposing stores off one bank ----
  constexpr int LDB = G + 1;
  if (gl < al) {
    for (int sr = 0; sr < 2 * ar; ++sr) {
      const int s = sr >= ar ? 1 : 0, r = sr - s * ar;
      Bs[(s * al + gl) * LDB + r] = valid ? Ain[sr * al + gl] : T(0);
    }
  }
  __syncwarp();
  // ---- M column of this lane: col[s * HR + k] = sum_l C[k][l] B[s][l][gl] ----
  T col[ROWS];
#pragma unroll
  for (int i = 0; i < ROWS; ++i) col[i] = T(0);
  const bool mine = gl < ar;
  for (int l = 0; l < al; ++l) {
    const T b0 = mine ? Bs[l * LDB + gl] : T(0);
    const T b1 = mine ? Bs[(al + l) * LDB + gl] : T(0);
#pragma unroll
    for (int k = 0; k < HR; ++k) {
      const T c = k < al ? C[k * G + l] : T(0);
      col[k] += c * b0;
      col[HR + k] += c * b1;
    }
  }
  T n2 = dot4<T, ROWS>(col, col);
  const T fro = group_sum_g<T, G>(n2);
  const T thr2 = Num<T>::floor_rel() * Num<T>::floor_rel() * fro;
  __syncwarp();  // every lane is done with Bs: R1 / R2 take its place
  for (int e = gl; e < ar * G; e += G) {
    R1[e] = T(0);
    R2[e] = T(0);
  }
  __syncwarp();
  // ---- pass 1: MGS in natural column order; live columns take steps 0, 1, ... ----
  int pos = -1;  // step of this lane's column (-1: dead or no column)
  int nk = 0;    // live columns so far (same in all lanes of the group)
  for (int t = 0; t < ar; ++t) {
    const T nt = __shfl_sync(FULL, n2, t, G);
    const bool live = nt > thr2;
    const T inv = live ? Num<T>::rsqrt(nt) : T(0);
    if (gl == t) {
#pragma unroll
      for (int i = 0; i < ROWS; ++i) col[i] *= inv;  // dead: the column becomes exactly zero
      store_q<T, ROWS>(qb, col);
      if (live) {
        pos = nk;
        R1[nk * G + t] = nt * inv;
      } else {
        n2 = T(0);
      }
    }
    __syncwarp();
    if (mine && gl > t && live) {
      T q[ROWS];
      load_q<T, ROWS>(q, qb);
      const T r = dot4<T, ROWS>(q, col);
#pragma unroll
      for (int i = 0; i < ROWS; ++i) col[i] -= r * q[i];
      n2 = dot4<T, ROWS>(col, col);
      R1[nk * G + gl] = r;
    }
    __syncwarp();
    nk += live ? 1 : 0;
  }
  const int rank = nk;
  const int rank_max = __reduce_max_sync(FULL, rank);
  // ---- pass 2: the same on Q1 (step order), R2[i][step] ----
  for (int i = 0; i < rank_max; ++i) {
    const bool on = i < rank;
    const unsigned who = __ballot_sync(FULL, pos == i);
    const unsigned gmask = (G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << ((threadIdx.x & 31) & ~(G - 1));
    const int src = on ? (__ffs(who & gmask) - 1) & (G - 1) : 0;
    const T m2 = dot4<T, ROWS>(col, col);
    const T ni = __shfl_sync(FULL, m2, src, G);
    const T inv = on ? Num<T>::rsqrt(ni) : T(0);
    if (on && gl == src) {
#pragma unroll
      for (int e = 0; e < ROWS; ++e) col[e] *= inv;
      store_q<T, ROWS>(qb, col);
      R2[i * G + i] = ni * inv;
    }
    __syncwarp();
    if (on && pos > i) {
      T q[ROWS];
      load_q<T, ROWS>(q, qb);
      const T r = dot4<T, ROWS>(q, col);
#pragma unroll
      for (int e = 0; e < ROWS; ++e) col[e] -= r * q[e];
      R2[i * G + pos] = r;
    }
    __syncwarp();
  }
  // ---- A_m[(s * al + k) * ar + slot]: live columns at their step, dead ones (zeros) behind ----
  {
    const unsigned gsh = (threadIdx.x & 31) & ~(G - 1);
    const unsigned deadmask = (__ballot_sync(FULL, mine && pos < 0) >> gsh) & (G == 32 ? 0xffffffffu : ((1u << G) - 1u));
    const int slot = pos >= 0 ? pos : rank + __popc(deadmask & ((1u << gl) - 1u));
    if (mine && valid) {
#pragma unroll
      for (int k = 0; k < HR; ++k) {
        if (k < al) {
          Aout[(size_t)k * ar + slot] = pos >= 0 ? col[k] : T(0);
          Aout[(size_t)(al + k) * ar + slot] = pos >= 0 ? col[HR + k] : T(0);
        }
      }
    }
  }
  // ---- carry C[i][j] = sum_{t >= i} R2[i][t] R1[t][j]  (rows >= rank zero) ----
  if (!last) {
    for (int i = 0; i < ar; ++i) {
      T acc = T(0);
      if (i < rank && mine)
        for (int t = i; t < rank; ++t) acc += R2[i * G + t] * R1[t * G + gl];
      C[i * G + gl] = acc;
    }
  }
  __syncwarp();
}

template <typename T, int G>
__global__ void __launch_bounds__(128) sweep_back_kernel(const __grid_constant__ SweepParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int IPW = 32 / G;  // images per warp
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int gl = lane & (G - 1), grp = lane / G;
  T* base = reinterpret_cast<T*>(smem_raw) + (size_t)(warp * IPW + grp) * sw_per_image(G);
  T* Bs = base;
  T* R1 = base;
  T* R2 = base + G * G;
  T* C = base + 2 * G * G + 2 * G;
  T* qb = C + G * G;
  const int L = P.L;
  for (int64_t w0 = (int64_t)blockIdx.x * wpb + warp; w0 * IPW < P.N; w0 += (int64_t)gridDim.x * wpb) {
    const int64_t img = w0 * IPW + grp;
    const bool valid = img < P.N;
    const int64_t im = valid ? img : 0;
    const T* tin = reinterpret_cast<const T*>(P.tmp_mps) + (size_t)im * P.mps_stride;
    T* out = reinterpret_cast<T*>(P.mps_out) + (size_t)im * P.mps_stride;
    if (gl == 0) C[0] = T(1);
    __syncwarp();
    for (int m = 0; m < L; ++m) {
      const int al = P.bonds[m], ar = P.bonds[m + 1];
      const T* Ain = tin + P.site_off[L - 1 - m];
      T* Aout = out + P.site_off[m];
      const bool last = m == L - 1;
      if (al <= 4) qr_site<T, 8, G>(Ain, Aout, valid, al, ar, Bs, C, R1, R2, qb, gl, last);
      else if (al <= 8) qr_site<T, 16, G>(Ain, Aout, valid, al, ar, Bs, C, R1, R2, qb, gl, last);
      else if (G > 8 && al <= 16) qr_site<T, (G > 8 ? 32 : 16), G>(Ain, Aout, valid, al, ar, Bs, C, R1, R2, qb, gl, last);
      else if (G > 16) qr_site<T, (G > 16 ? 64 : 16), G>(Ain, Aout, valid, al, ar, Bs, C, R1, R2, qb, gl, last);
    }
    // singular values / sweep counts the right-to-left sweep saw: row n (bond n + 1) is step
    // L - 2 - n of the mirrored chain; row L - 1 (the norm of the truncated image) stays
    if (P.sv_out && valid) {
      const T* si = reinterpret_cast<const T*>(P.tmp_sv) + (size_t)img * L * P.sv_width;
      T* so = reinterpret_cast<T*>(P.sv_out) + (size_t)img * L * P.sv_width;
      for (int e = gl; e < L * P.sv_width; e += G) {
        const int n = e / P.sv_width, k = e % P.sv_width;
        const int src = n == L - 1 ? n : L - 2 - n;
        so[e] = si[src * P.sv_width + k];
      }
    }
    if (P.sw_out && valid)
      for (int n = gl; n < L; n += G) {
        const int src = n == L - 1 ? n : L - 2 - n;
        P.sw_out[(size_t)img * L + n] = P.tmp_sw[(size_t)img * L + src];
      }
    __syncwarp();
  }
}

}  // namespace sweep
}  // namespace oc
