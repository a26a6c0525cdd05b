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

constexpr int kSwLd = 32;                        // row stride of the per-warp square buffers
constexpr int kSwPerWarp = 2048 + 3 * 1024 + 64; // Bs | C | R1 | R2 | q   (elements of T)

// One site with ROWS = 2 * HR register rows per column (HR >= a_m).
template <typename T, int ROWS>
__device__ __forceinline__ void qr_site(const T* __restrict__ Ain, T* __restrict__ Aout, int al, int ar, T* Bs,
                                        T* C, T* R1, T* R2, T* qb, int lane, bool last) {
  constexpr int HR = ROWS / 2;
  // ---- stage B_m[s][l][r] = A'[s][r][l] as Bs[(s * al + l) * 32 + r] ----
  for (int e = lane; e < 2 * ar * al; e += 32) {
    const int l = e % al, sr = e / al, r = sr % ar, s = sr / ar;
    Bs[(s * al + l) * kSwLd + r] = Ain[e];
  }
  __syncwarp();
  // ---- M column of this lane: col[s * HR + k] = sum_l C[k][l] B[s][l][lane] ----
  T col[ROWS];
#pragma unroll
  for (int i = 0; i < ROWS; ++i) col[i] = T(0);
  const bool mine = lane < ar;
  for (int l = 0; l < al; ++l) {
    const T b0 = mine ? Bs[l * kSwLd + lane] : T(0);
    const T b1 = mine ? Bs[(al + l) * kSwLd + lane] : T(0);
#pragma unroll
    for (int k = 0; k < HR; ++k) {
      const T c = k < al ? C[k * kSwLd + l] : T(0);
      col[k] += c * b0;
      col[HR + k] += c * b1;
    }
  }
  T n2 = T(0);
#pragma unroll
  for (int i = 0; i < ROWS; ++i) n2 += col[i] * col[i];
  const T fro = warp_sum(n2);
  const T thr2 = Num<T>::floor_rel() * Num<T>::floor_rel() * fro;
  // zero-fill R1 / R2 rows this site can touch
  for (int e = lane; e < ar * kSwLd; e += 32) {
    R1[e] = T(0);
    R2[e] = T(0);
  }
  __syncwarp();
  // ---- pass 1: MGS in natural column order; live columns take steps 0, 1, ... ----
  int pos = -1;  // step of this lane's column (-1: dead or no column)
  int nk = 0;
  for (int t = 0; t < ar; ++t) {
    const T nt = __shfl_sync(FULL, n2, t);
    if (!(nt > thr2)) {  // warp-uniform: dead column
      if (lane == t) {
#pragma unroll
        for (int i = 0; i < ROWS; ++i) col[i] = T(0);
        n2 = T(0);
      }
      continue;
    }
    const T inv = Num<T>::rsqrt(nt);
    if (lane == t) {
#pragma unroll
      for (int i = 0; i < ROWS; ++i) {
        col[i] *= inv;
        qb[i] = col[i];
      }
      pos = nk;
      R1[nk * kSwLd + t] = nt * inv;
    }
    __syncwarp();
    if (mine && lane > t) {
      T r = T(0);
#pragma unroll
      for (int i = 0; i < ROWS; ++i) r += qb[i] * col[i];
      T acc = T(0);
#pragma unroll
      for (int i = 0; i < ROWS; ++i) {
        col[i] -= r * qb[i];
        acc += col[i] * col[i];
      }
      n2 = acc;
      R1[nk * kSwLd + lane] = r;
    }
    __syncwarp();
    ++nk;
  }
  const int rank = nk;
  // ---- pass 2: the same on Q1 (step order), R2[i][step] ----
  for (int i = 0; i < rank; ++i) {
    const unsigned who = __ballot_sync(FULL, pos == i);
    const int src = __ffs(who) - 1;
    T m2 = T(0);
#pragma unroll
    for (int e = 0; e < ROWS; ++e) m2 += col[e] * col[e];
    const T ni = __shfl_sync(FULL, m2, src);
    const T inv = Num<T>::rsqrt(ni);
    if (lane == src) {
#pragma unroll
      for (int e = 0; e < ROWS; ++e) {
        col[e] *= inv;
        qb[e] = col[e];
      }
      R2[i * kSwLd + i] = ni * inv;
    }
    __syncwarp();
    if (pos > i) {
      T r = T(0);
#pragma unroll
      for (int e = 0; e < ROWS; ++e) r += qb[e] * col[e];
#pragma unroll
      for (int e = 0; e < ROWS; ++e) col[e] -= r * qb[e];
      R2[i * kSwLd + pos] = r;
    }
    __syncwarp();
  }
  // ---- A_m[(s * al + k) * ar + slot]: live columns at their step, dead ones (zeros) behind ----
  {
    const unsigned deadmask = __ballot_sync(FULL, mine && pos < 0);
    const int slot = pos >= 0 ? pos : rank + __popc(deadmask & ((1u << lane) - 1u));
    if (mine) {
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
        for (int t = i; t < rank; ++t) acc += R2[i * kSwLd + t] * R1[t * kSwLd + lane];
      if (lane < kSwLd) C[i * kSwLd + lane] = acc;
    }
  }
  __syncwarp();
}

template <typename T>
__global__ void __launch_bounds__(128) sweep_back_kernel(const __grid_constant__ SweepParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  T* base = reinterpret_cast<T*>(smem_raw) + (size_t)warp * kSwPerWarp;
  T* Bs = base;
  T* C = base + 2048;
  T* R1 = C + 1024;
  T* R2 = R1 + 1024;
  T* qb = R2 + 1024;
  const int L = P.L;
  for (int64_t img = (int64_t)blockIdx.x * wpb + warp; img < P.N; img += (int64_t)gridDim.x * wpb) {
    const T* tin = reinterpret_cast<const T*>(P.tmp_mps) + (size_t)img * P.mps_stride;
    T* out = reinterpret_cast<T*>(P.mps_out) + (size_t)img * P.mps_stride;
    if (lane == 0) C[0] = T(1);
    __syncwarp();
    for (int m = 0; m < L; ++m) {
      const int al = P.bonds[m], ar = P.bonds[m + 1];
      const T* Ain = tin + P.site_off[L - 1 - m];
      T* Aout = out + P.site_off[m];
      const bool last = m == L - 1;
      if (al <= 4) qr_site<T, 8>(Ain, Aout, al, ar, Bs, C, R1, R2, qb, lane, last);
      else if (al <= 8) qr_site<T, 16>(Ain, Aout, al, ar, Bs, C, R1, R2, qb, lane, last);
      else if (al <= 16) qr_site<T, 32>(Ain, Aout, al, ar, Bs, C, R1, R2, qb, lane, last);
      else qr_site<T, 64>(Ain, Aout, al, ar, Bs, C, R1, R2, qb, lane, last);
    }
    // singular values / sweep counts the right-to-left sweep saw: row n (bond n + 1) is step
    // L - 2 - n of the mirrored chain; row L - 1 (the norm of the truncated image) stays
    if (P.sv_out) {
      const T* si = reinterpret_cast<const T*>(P.tmp_sv) + (size_t)img * L * P.sv_width;
      T* so = reinterpret_cast<T*>(P.sv_out) + (size_t)img * L * P.sv_width;
      for (int e = lane; e < L * P.sv_width; e += 32) {
        const int n = e / P.sv_width, k = e % P.sv_width;
        const int src = n == L - 1 ? n : L - 2 - n;
        so[e] = si[src * P.sv_width + k];
      }
    }
    if (P.sw_out && lane < L) {
      const int src = lane == L - 1 ? lane : L - 2 - lane;
      P.sw_out[(size_t)img * L + lane] = P.tmp_sw[(size_t)img * L + src];
    }
    __syncwarp();
  }
}

}  // namespace sweep
}  // namespace oc
