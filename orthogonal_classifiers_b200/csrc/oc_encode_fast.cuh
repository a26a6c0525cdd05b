// Specialised encoder for L = 10 (512 < n_pix <= 1024), fp32, one image per
// warp, the vectors being orthogonalised live in REGISTERS.
//
// Step n works on M_n (P = 2 B rows, Q cols), B = b_n, Q = 2^(9-n).  The NV
// vectors (rows when P < Q, columns otherwise) are laid out over the warp as
// NV/2 "slots" of LP = 64/NV lanes; a slot owns the two vectors at line
// positions (2k, 2k+1) — register sets T and S, each lane holding EL = LEN/LP
// consecutive elements of both.  One-sided Jacobi with the odd-even ordering:
//   round A: rotate (T_k, S_k)            — no communication at all
//   round B: rotate (S_k, T_{k+1})        — T moves one slot down and back
// each rotation also swaps the two positions, so after NV rounds every pair of
// vectors has met exactly once (a sweep).  Dot products reduce over the LP
// lanes of a slot with log2(LP) xor-shuffles.
//
// Row mode additionally carries each vector's row of U^T (the accumulated
// rotations, = the site tensor A_n); column mode gets A_n by normalising the
// converged columns and the next Psi from one small U^T M product.
#pragma once
#include "oc_generic.cuh"

namespace oc {

constexpr float kTol2 = 1e-12f;     // (1e-6)^2: rotate if (a.b)^2 > tol^2 |a|^2|b|^2
constexpr float kQStop2 = 1e-7f;    // (3.2e-4)^2: a sweep whose largest rotation is below this is the last
constexpr float kFloor2 = 1e-12f;   // (1e-6)^2 relative to ||M||_F^2: numerically zero vectors
constexpr int kMaxSweeps = 14;

template <int B, int Q>
struct StepCfg {
  static constexpr int P = 2 * B;
  static constexpr bool ROW = P < Q;
  static constexpr int NV = ROW ? P : Q;
  static constexpr int LEN = ROW ? Q : P;
  static constexpr int LP = NV >= 2 ? 64 / NV : 32;
  static constexpr int EL = LEN >= LP ? LEN / LP : 1;
  static constexpr int UEL = ROW ? (NV >= LP ? NV / LP : 1) : 1;
};

// rotate the pair (a, b) held in registers; afterwards a holds the vector for
// the FIRST line position (b') and b the one for the second (a'): rotate+swap.
template <int EL, int UEL, int LP, bool ROW>
__device__ __forceinline__ void rotate_pair(float (&a)[EL], float (&b)[EL], float (&ua)[UEL],
                                            float (&ub)[UEL], float thr2, bool idle, bool& again) {
  float aa = 0.f, bb = 0.f, ab = 0.f;
#pragma unroll
  for (int e = 0; e < EL; ++e) {
    aa = fmaf(a[e], a[e], aa);
    bb = fmaf(b[e], b[e], bb);
    ab = fmaf(a[e], b[e], ab);
  }
#pragma unroll
  for (int o = 1; o < LP; o <<= 1) {
    aa += __shfl_xor_sync(FULL, aa, o);
    bb += __shfl_xor_sync(FULL, bb, o);
    ab += __shfl_xor_sync(FULL, ab, o);
  }
  const float d = aa * bb;
  const float ab2 = ab * ab;
  const bool rot = (ab2 > kTol2 * d) && (aa > thr2) && (bb > thr2) && !idle;
  again = again || (rot && ab2 > kQStop2 * d);
  float c = 1.f, s = 0.f;
  if (rot) {
    const float zeta = __fdividef(bb - aa, 2.f * ab);
    float t = __fdividef(1.f, fabsf(zeta) + sqrtf(fmaf(zeta, zeta, 1.f)));
    t = zeta < 0.f ? -t : t;
    const float h = fmaf(t, t, 1.f);
    float r = rsqrtf(h);
    r = r * fmaf(-0.5f * h, r * r, 1.5f);  // one Newton step: c^2 + s^2 = 1 to ~1 ulp
    c = r;
    s = r * t;
  }
  if (idle) {  // keep both vectors where they are (the second one picks up a sign)
    c = 0.f;
    s = 1.f;
  }
#pragma unroll
  for (int e = 0; e < EL; ++e) {
    const float x = a[e], y = b[e];
    a[e] = fmaf(s, x, c * y);
    b[e] = fmaf(c, x, -s * y);
  }
  if (ROW) {
#pragma unroll
    for (int e = 0; e < UEL; ++e) {
      const float x = ua[e], y = ub[e];
      ua[e] = fmaf(s, x, c * y);
      ub[e] = fmaf(c, x, -s * y);
    }
  }
}

template <int NV, int EL, int UEL, bool ROW>
__device__ __forceinline__ int jacobi_regs(float (&T)[EL], float (&S)[EL], float (&UT)[UEL],
                                           float (&US)[UEL], float thr2, int lane) {
  constexpr int LP = 64 / NV;
  constexpr int M = NV / 2;
  const int slot = lane / LP;
  const int src_dn = (lane + LP) & 31;
  const int src_up = (lane - LP) & 31;
  int sweeps = 0;
  while (sweeps < kMaxSweeps) {
    bool again = false;
#pragma unroll 1
    for (int rr = 0; rr < M; ++rr) {
      rotate_pair<EL, UEL, LP, ROW>(T, S, UT, US, thr2, false, again);
      if (NV > 2) {
        float R[EL], UR[UEL];
#pragma unroll
        for (int e = 0; e < EL; ++e) R[e] = __shfl_sync(FULL, T[e], src_dn);
        if (ROW) {
#pragma unroll
          for (int e = 0; e < UEL; ++e) UR[e] = __shfl_sync(FULL, UT[e], src_dn);
        }
        rotate_pair<EL, UEL, LP, ROW>(S, R, US, UR, thr2, slot == M - 1, again);
#pragma unroll
        for (int e = 0; e < EL; ++e) T[e] = __shfl_sync(FULL, R[e], src_up);
        if (ROW) {
#pragma unroll
          for (int e = 0; e < UEL; ++e) UT[e] = __shfl_sync(FULL, UR[e], src_up);
        }
      }
    }
    ++sweeps;
    if (NV == 2) break;  // one rotation orthogonalises two vectors exactly
    if (!__any_sync(FULL, again)) break;
  }
  return sweeps;
}

// One step of the chain.  in:  Psi_n, either the image in global memory
// (FIRST) or smem `Pin` laid out [B][2Q].  out: Psi_{n+1} [R][Q] in smem; the
// function returns the buffer that holds it (Pin for column mode, Pout for row
// mode).  Site tensor A_n goes straight to global memory.
template <int B, int Q, int R, bool FIRST>
__device__ __forceinline__ float* jstep(const EncodeParams& P, int n, int64_t img, float* Pin,
                                        float* Pout, float* mps, int lane) {
  using C = StepCfg<B, Q>;
  constexpr int NV = C::NV, LEN = C::LEN, LP = C::LP, EL = C::EL, UEL = C::UEL;
  constexpr bool ROW = C::ROW;
  const int b_rt = P.bonds[n];      // runtime bonds of the output layout (b_n(D) <= B)
  const int r_rt = P.bonds[n + 1];
  float* A = mps + P.site_off[n];
  float* sv = P.svals_out ? reinterpret_cast<float*>(P.svals_out) + ((size_t)img * P.L + n) * P.sv_width : nullptr;

  if constexpr (NV == 1) {
    // last step: M (P x 1) -> A = M / |M|; the 1x1 S.V is dropped (unit norm)
    float v = lane < LEN ? Pin[(lane % B) * 2 * Q + (lane / B) * Q] : 0.f;
    float nn = warp_sum(v * v);
    const float inv = nn > 0.f ? rsqrtf(nn) : 0.f;
    const int s = lane / B, l = lane % B;
    if (lane < LEN && l < b_rt) A[(s * b_rt + l) * r_rt] = v * inv;
    if (sv) {
      for (int k = lane; k < P.sv_width; k += 32) sv[k] = k == 0 ? sqrtf(nn) : 0.f;
    }
    if (P.sweeps_out && lane == 0) P.sweeps_out[(size_t)img * P.L + n] = 0;
    return Pin;
  } else {
    const int slot = lane / LP, sub = lane % LP;
    float T[EL], S[EL], UT[UEL], US[UEL];
    // ---- load the two vectors of this slot ----
    if constexpr (FIRST) {
      // B = 1: T = pixels [0, 512), S = pixels [512, 1024); lane holds 16 of each
      const char* src = reinterpret_cast<const char*>(P.images);
      const int esz = P.in_dtype == 0 ? 4 : (P.in_dtype == 1 ? 8 : 1);
      const void* row = src + (size_t)img * P.ld * esz;
      const bool vec = P.in_dtype == 0 && ((P.ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
#pragma unroll
      for (int e = 0; e < EL; e += 4) {
        const int iT = sub * EL + e, iS = Q + sub * EL + e;
        if (vec && iT + 3 < P.n_pix) {
          float4 v = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(row) + iT));
          T[e] = v.x; T[e + 1] = v.y; T[e + 2] = v.z; T[e + 3] = v.w;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) T[e + j] = iT + j < P.n_pix ? load_pixel<float>(row, P.in_dtype, iT + j) : 0.f;
        }
        if (vec && iS + 3 < P.n_pix) {
          float4 v = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(row) + iS));
          S[e] = v.x; S[e + 1] = v.y; S[e + 2] = v.z; S[e + 3] = v.w;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) S[e + j] = iS + j < P.n_pix ? load_pixel<float>(row, P.in_dtype, iS + j) : 0.f;
        }
      }
    } else if constexpr (ROW) {
      // vector j = 2l + s is the chunk [l][s*Q .. s*Q+Q) of Psi_n; slot k = l
#pragma unroll
      for (int e = 0; e < EL; ++e) {
        const int idx = sub * EL + e;
        T[e] = idx < LEN ? Pin[slot * 2 * Q + idx] : 0.f;
        S[e] = idx < LEN ? Pin[slot * 2 * Q + Q + idx] : 0.f;
      }
    } else {
      // vector c = column c of M; element e' = s*B + l lives at Psi_n[l][s*Q + c]
#pragma unroll
      for (int e = 0; e < EL; ++e) {
        const int idx = sub * EL + e;
        const int s = idx / B, l = idx % B;
        const bool ok = idx < LEN;
        T[e] = ok ? Pin[l * 2 * Q + s * Q + 2 * slot] : 0.f;
        S[e] = ok ? Pin[l * 2 * Q + s * Q + 2 * slot + 1] : 0.f;
      }
    }
#pragma unroll
    for (int e = 0; e < UEL; ++e) {
      const int jj = sub * UEL + e;
      UT[e] = (ROW && jj == 2 * slot) ? 1.f : 0.f;
      US[e] = (ROW && jj == 2 * slot + 1) ? 1.f : 0.f;
    }
    // ---- Frobenius norm -> zero threshold ----
    float fro = 0.f;
#pragma unroll
    for (int e = 0; e < EL; ++e) fro = fmaf(T[e], T[e], fmaf(S[e], S[e], fro));
    fro = warp_sum(fro);
    const float thr2 = kFloor2 * fro;

    int sweeps = 0;
    if (fro > 0.f) sweeps = jacobi_regs<NV, EL, UEL, ROW>(T, S, UT, US, thr2, lane);

    // ---- norms, stable descending rank ----
    float nT = 0.f, nS = 0.f;
#pragma unroll
    for (int e = 0; e < EL; ++e) {
      nT = fmaf(T[e], T[e], nT);
      nS = fmaf(S[e], S[e], nS);
    }
#pragma unroll
    for (int o = 1; o < LP; o <<= 1) {
      nT += __shfl_xor_sync(FULL, nT, o);
      nS += __shfl_xor_sync(FULL, nS, o);
    }
    int rT = 0, rS = 0;
    const int pT = 2 * slot, pS = 2 * slot + 1;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const float v = __shfl_sync(FULL, (j & 1) ? nS : nT, (j >> 1) * LP);
      rT += (v > nT) || (v == nT && j < pT);
      rS += (v > nS) || (v == nS && j < pS);
    }
    const bool kT = nT > thr2 && rT < r_rt;
    const bool kS = nS > thr2 && rS < r_rt;
    if (sv) {
      if (sub == 0) {
        if (rT < P.sv_width) sv[rT] = nT > thr2 ? sqrtf(nT) : 0.f;
        if (rS < P.sv_width) sv[rS] = nS > thr2 ? sqrtf(nS) : 0.f;
      }
      for (int k = NV + lane; k < P.sv_width; k += 32) sv[k] = 0.f;
    }
    if (P.sweeps_out && lane == 0) P.sweeps_out[(size_t)img * P.L + n] = sweeps;

    if constexpr (ROW) {
      // A_n[(s,l), rank] = U-row element jj = 2l + s
#pragma unroll
      for (int e = 0; e < UEL; ++e) {
        const int jj = sub * UEL + e;
        const int l = jj >> 1, s = jj & 1;
        if (jj < NV && l < b_rt) {
          if (rT < r_rt) A[(s * b_rt + l) * r_rt + rT] = kT ? UT[e] : 0.f;
          if (rS < r_rt) A[(s * b_rt + l) * r_rt + rS] = kS ? US[e] : 0.f;
        }
      }
      // next Psi: row rank <- vector (zero when dropped)
#pragma unroll
      for (int e = 0; e < EL; ++e) {
        const int idx = sub * EL + e;
        if (idx < LEN) {
          if (rT < R) Pout[rT * Q + idx] = kT ? T[e] : 0.f;
          if (rS < R) Pout[rS * Q + idx] = kS ? S[e] : 0.f;
        }
      }
      __syncwarp();
      return Pout;
    } else {
      // normalise -> U columns; stage U as [e'][R] in Pout; A_n to global
      const float iT = kT ? rsqrtf(nT) : 0.f;
      const float iS = kS ? rsqrtf(nS) : 0.f;
#pragma unroll
      for (int e = 0; e < EL; ++e) {
        const int idx = sub * EL + e;
        if (idx < LEN) {
          const int s = idx / B, l = idx % B;
          const float uT = T[e] * iT, uS = S[e] * iS;
          if (rT < R) Pout[idx * R + rT] = uT;
          if (rS < R) Pout[idx * R + rS] = uS;
          if (l < b_rt) {
            if (rT < r_rt) A[(s * b_rt + l) * r_rt + rT] = uT;
            if (rS < r_rt) A[(s * b_rt + l) * r_rt + rS] = uS;
          }
        }
      }
      __syncwarp();
      // Psi_{n+1}[r][c] = sum_e U[e][r] M[e][c],  M[e = s*B + l][c] = Pin[l][s*Q + c]
      constexpr int NG = Q >= 32 ? 1 : 32 / Q;          // lane groups sharing a column
      constexpr int RPL = R >= NG ? R / NG : 1;          // rows per lane
      const int c = lane % Q, g = lane / Q;
      const bool active = (g * RPL) < R;
      float acc[RPL];
#pragma unroll
      for (int i = 0; i < RPL; ++i) acc[i] = 0.f;
      if (active) {
#pragma unroll 4
        for (int e = 0; e < LEN; ++e) {
          const int s = e / B, l = e % B;
          const float m = Pin[l * 2 * Q + s * Q + c];
          const float* u = Pout + e * R + g * RPL;
#pragma unroll
          for (int i = 0; i < RPL; ++i) acc[i] = fmaf(u[i], m, acc[i]);
        }
      }
      __syncwarp();
      if (active) {
#pragma unroll
        for (int i = 0; i < RPL; ++i) Pin[(g * RPL + i) * Q + c] = acc[i];
      }
      __syncwarp();
      return Pin;
    }
  }
}

template <int DPAD>
__global__ void __launch_bounds__(128) encode_fast_kernel(const __grid_constant__ EncodeParams P);

// D_pad = 32 chain: b = 1,2,4,8,16,32,16,8,4,2,1
template <>
__global__ void __launch_bounds__(128) encode_fast_kernel<32>(const __grid_constant__ EncodeParams P) {
  extern __shared__ __align__(16) float smem_f[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  float* b0 = smem_f + (size_t)warp * 2048;
  float* b1 = b0 + 1024;
  for (int64_t img = (int64_t)blockIdx.x * wpb + warp; img < P.N; img += (int64_t)gridDim.x * wpb) {
    float* mps = reinterpret_cast<float*>(P.mps_out) + (size_t)img * P.mps_stride;
    float* cur;
    cur = jstep<1, 512, 2, true>(P, 0, img, nullptr, b0, mps, lane);
    cur = jstep<2, 256, 4, false>(P, 1, img, cur, cur == b0 ? b1 : b0, mps, lane);
    cur = jstep<4, 128, 8, false>(P, 2, img, cur, cur == b0 ? b1 : b0, mps, lane);
    cur = jstep<8, 64, 16, false>(P, 3, img, cur, cur == b0 ? b1 : b0, mps, lane);
    cur = jstep<16, 32, 32, false>(P, 4, img, cur, cur == b0 ? b1 : b0, mps, lane);
    cur = jstep<32, 16, 16, false>(P, 5, img, cur, cur == b0 ? b1 : b0, mps, lane);
    cur = jstep<16, 8, 8, false>(P, 6, img, cur, cur == b0 ? b1 : b0, mps, lane);
    cur = jstep<8, 4, 4, false>(P, 7, img, cur, cur == b0 ? b1 : b0, mps, lane);
    cur = jstep<4, 2, 2, false>(P, 8, img, cur, cur == b0 ? b1 : b0, mps, lane);
    cur = jstep<2, 1, 1, false>(P, 9, img, cur, cur == b0 ? b1 : b0, mps, lane);
    __syncwarp();
  }
}

inline bool encode_fast_supported(const EncodeParams& P) { return P.L == 10; }

inline int encode_fast_launch(const EncodeParams& P, int sms, cudaStream_t st) {
  const int warps = 4;
  const size_t smem = (size_t)warps * 2048 * sizeof(float);
  const int64_t need = (P.N + warps - 1) / warps;
  const int grid = (int)(need < (int64_t)sms * 12 ? need : (int64_t)sms * 12);
  encode_fast_kernel<32><<<grid, warps * 32, smem, st>>>(P);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

}  // namespace oc
