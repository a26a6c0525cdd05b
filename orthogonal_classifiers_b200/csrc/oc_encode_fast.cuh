// Specialised encoder for L = 10 (512 < n_pix <= 1024), fp32, one image per
// warp; the vectors being orthogonalised live in REGISTERS.
//
// Step n works on M_n (P = 2B rows, Q cols), B = b_n, Q = 2^(9-n).  Its NV
// vectors (rows when P < Q, columns otherwise; length LEN) are spread over the
// warp as "slots" of LP lanes; a slot owns the two vectors at line positions
// (2k, 2k+1) — register sets T and S, each lane holding EL consecutive elements
// of both (EL = 16 for the big steps, 4 for the small tail steps).
// One-sided (Hestenes) Jacobi with the odd-even ordering:
//   round A: rotate (T_k, S_k)        — no communication
//   round B: rotate (S_k, T_{k+1})    — T moves one slot down and back
// every rotation also swaps the two line positions, so after NV rounds every
// pair has met once (one sweep).  Squared norms are carried with the vectors
// and updated by the rotation (refreshed from the data every sweep and
// whenever an update cancels), so a round costs ONE dot product.  All
// element-wise arithmetic is packed fma.rn.f32x2 (SASS FFMA2).
//
// ONE step body, parametrised at run time by a small per-step table, serves
// all ten steps and every D: the hot loop is ~1 K instructions and stays in the
// instruction cache (the first, fully unrolled version was 246 KB of SASS and
// stalled on instruction fetch — profiles/encode_r1a.md).
#pragma once
#include "oc_generic.cuh"

namespace oc {

typedef unsigned long long u64;

constexpr float kTol2 = 1e-12f;    // rotate if (a.b)^2 > (1e-6)^2 |a|^2 |b|^2
constexpr float kQStop2 = 1e-7f;   // a sweep whose largest rotation had cos^2 below this is the last
constexpr float kFloor2 = 1e-12f;  // vectors below 1e-6 ||M||_F are numerically zero
constexpr float kCancel = 0.05f;   // norm update lost > 95 %: recompute it from the data
constexpr int kMaxSweeps = 14;

__device__ __forceinline__ u64 pack2(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
  u64 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float hsum2(u64 v) {
  float lo, hi;
  unpack2(v, lo, hi);
  return lo + hi;
}
__device__ __forceinline__ u64 shfl2(u64 v, int src) {
  float lo, hi;
  unpack2(v, lo, hi);
  lo = __shfl_sync(FULL, lo, src);
  hi = __shfl_sync(FULL, hi, src);
  return pack2(lo, hi);
}

struct FastStep {
  int logB, logQ;  // Psi_n is [B][2Q] in shared memory (padded to powers of two)
  int row;         // 1: the vectors are the rows of M_n
  int NV, LEN;     // number of vectors / their length
  int logLP;       // lanes per slot
  int M;           // active slots = NV / 2
  int el16;        // 1: EL = 16 instance, 0: EL = 4
  int R;           // rows of Psi_{n+1} kept in shared memory (padded)
  int rpl;         // rows per lane of the U^T M product (32, 8, 2 or 1)
};
struct FastPlan {
  FastStep st[OC_MAX_L];
};

template <int E2>
__device__ __forceinline__ float group_dot(const u64 (&a)[E2], const u64 (&b)[E2], int LP) {
  u64 acc = 0ull;
#pragma unroll
  for (int e = 0; e < E2; ++e) acc = fma2(a[e], b[e], acc);
  float v = hsum2(acc);
  for (int o = 1; o < LP; o <<= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// rotate (a, b) with cached squared norms (na, nb); afterwards a/na hold the
// vector of the FIRST line position (b') and b/nb the second (a').
template <int E2>
__device__ __forceinline__ void rotate_pair(u64 (&a)[E2], u64 (&b)[E2], u64 (&ua)[2], u64 (&ub)[2],
                                            float& na, float& nb, float thr2, bool idle, bool row,
                                            int LP, bool& again) {
  const float ab = group_dot<E2>(a, b, LP);
  const float aa = na, bb = nb;
  const float d = aa * bb;
  const float ab2 = ab * ab;
  const bool rot = (ab2 > kTol2 * d) && (aa > thr2) && (bb > thr2) && !idle;
  again = again || (rot && ab2 > kQStop2 * d);
  float c = 1.f, s = 0.f, t = 0.f;
  if (rot) {
    const float zeta = __fdividef(bb - aa, 2.f * ab);
    t = __fdividef(1.f, fabsf(zeta) + sqrtf(fmaf(zeta, zeta, 1.f)));
    t = zeta < 0.f ? -t : t;
    const float h = fmaf(t, t, 1.f);
    float r = rsqrtf(h);
    r = r * fmaf(-0.5f * h, r * r, 1.5f);  // Newton step: c^2 + s^2 = 1 to ~1 ulp
    c = r;
    s = r * t;
  }
  if (idle) {  // both vectors stay where they are (the second picks up a sign)
    c = 0.f;
    s = 1.f;
  }
  // |a'|^2 = aa - t ab, |b'|^2 = bb + t ab; positions swap
  float n_first = fmaf(t, ab, bb);
  float n_second = fmaf(-t, ab, aa);
  if (idle) {
    n_first = aa;
    n_second = bb;
  }
  const bool lost = rot && (n_second < kCancel * aa || n_first < kCancel * bb);
  const u64 c2 = pack2(c, c), s2 = pack2(s, s), ns2 = pack2(-s, -s);
#pragma unroll
  for (int e = 0; e < E2; ++e) {
    const u64 x = a[e], y = b[e];
    a[e] = fma2(s2, x, mul2(c2, y));
    b[e] = fma2(c2, x, mul2(ns2, y));
  }
  if (row) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const u64 x = ua[e], y = ub[e];
      ua[e] = fma2(s2, x, mul2(c2, y));
      ub[e] = fma2(c2, x, mul2(ns2, y));
    }
  }
  if (__any_sync(FULL, lost)) {  // warp-uniform, rare after the first sweeps
    n_first = group_dot<E2>(a, a, LP);
    n_second = group_dot<E2>(b, b, LP);
  }
  na = n_first;
  nb = n_second;
}

template <int E2>
__device__ __forceinline__ int jacobi_regs(u64 (&T)[E2], u64 (&S)[E2], u64 (&UT)[2], u64 (&US)[2],
                                           float& nT, float& nS, float thr2, int lane, int logLP,
                                           int M, bool row) {
  const int LP = 1 << logLP;
  const int slot = lane >> logLP;
  const int src_dn = (lane + LP) & 31;
  const int src_up = (lane - LP) & 31;
  const bool idleB = slot >= M - 1;
  int sweeps = 0;
#pragma unroll 1
  while (sweeps < kMaxSweeps) {
    bool again = false;
    nT = group_dot<E2>(T, T, LP);
    nS = group_dot<E2>(S, S, LP);
#pragma unroll 1
    for (int rr = 0; rr < M; ++rr) {
      rotate_pair<E2>(T, S, UT, US, nT, nS, thr2, false, row, LP, again);
      if (M > 1) {
        u64 R[E2], UR[2];
#pragma unroll
        for (int e = 0; e < E2; ++e) R[e] = shfl2(T[e], src_dn);
        float nR = __shfl_sync(FULL, nT, src_dn);
        if (row) {
          UR[0] = shfl2(UT[0], src_dn);
          UR[1] = shfl2(UT[1], src_dn);
        } else {
          UR[0] = UR[1] = 0ull;
        }
        rotate_pair<E2>(S, R, US, UR, nS, nR, thr2, idleB, row, LP, again);
#pragma unroll
        for (int e = 0; e < E2; ++e) T[e] = shfl2(R[e], src_up);
        nT = __shfl_sync(FULL, nR, src_up);
        if (row) {
          UT[0] = shfl2(UR[0], src_up);
          UT[1] = shfl2(UR[1], src_up);
        }
      }
    }
    ++sweeps;
    if (M == 1) break;  // one rotation orthogonalises two vectors exactly
    if (!__any_sync(FULL, again)) break;
  }
  return sweeps;
}

// Psi_{n+1}[r][c] = sum_e U[e][r] M[e][c];  U staged as [e][R] in `U`,
// M[e = s*B + l][c] = Pin[l][s*Q + c].  Lane owns column c and RPL rows.
template <int RPL>
__device__ __forceinline__ void utm_product(float* Pin, const float* U, int LEN, int logB, int logQ,
                                            int R, int lane) {
  const int Q = 1 << logQ, B = 1 << logB;
  const int c = lane & (Q - 1), g = lane >> logQ;
  const bool active = g * RPL < R;
  float acc[RPL];
#pragma unroll
  for (int i = 0; i < RPL; ++i) acc[i] = 0.f;
  if (active) {
#pragma unroll 2
    for (int e = 0; e < LEN; ++e) {
      const int s = e >> logB, l = e & (B - 1);
      const float m = Pin[l * 2 * Q + s * Q + c];
      const float* u = U + e * R + g * RPL;
      if (RPL >= 4) {
#pragma unroll
        for (int i = 0; i < RPL; i += 4) {
          const float4 uv = *reinterpret_cast<const float4*>(u + i);
          acc[i] = fmaf(uv.x, m, acc[i]);
          acc[i + 1] = fmaf(uv.y, m, acc[i + 1]);
          acc[i + 2] = fmaf(uv.z, m, acc[i + 2]);
          acc[i + 3] = fmaf(uv.w, m, acc[i + 3]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < RPL; ++i) acc[i] = fmaf(u[i], m, acc[i]);
      }
    }
  }
  __syncwarp();
  if (active) {
#pragma unroll
    for (int i = 0; i < RPL; ++i) Pin[(g * RPL + i) * Q + c] = acc[i];
  }
  __syncwarp();
}

// One step of the chain (see file header).  Returns the buffer holding Psi_{n+1}.
template <int EL>
__device__ __forceinline__ float* run_step(const EncodeParams& P, const FastStep& st, int n,
                                           int64_t img, float* Pin, float* Pout, float* mps, int lane) {
  constexpr int E2 = EL / 2;
  const int logLP = st.logLP, LP = 1 << logLP;
  const int slot = lane >> logLP, sub = lane & (LP - 1);
  const int B = 1 << st.logB, Q = 1 << st.logQ;
  const int NV = st.NV, LEN = st.LEN, R = st.R;
  const bool row = st.row != 0;
  const bool act = slot < st.M;
  const int b_rt = P.bonds[n], r_rt = P.bonds[n + 1];
  float* A = mps + P.site_off[n];
  float* sv = P.svals_out ? reinterpret_cast<float*>(P.svals_out) + ((size_t)img * P.L + n) * P.sv_width : nullptr;

  u64 T[E2], S[E2], UT[2], US[2];
  const int idx0 = sub * EL;  // first element index held by this lane
  // ---- load the two vectors of this slot ----
  if (n == 0) {
    // B = 1: T = pixels [0, 512), S = pixels [512, 1024)
    const char* src = reinterpret_cast<const char*>(P.images);
    const int esz = P.in_dtype == 0 ? 4 : (P.in_dtype == 1 ? 8 : 1);
    const void* rowp = src + (size_t)img * P.ld * esz;
    const bool vec = P.in_dtype == 0 && ((P.ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
#pragma unroll
    for (int e = 0; e < EL; e += 4) {
      float t[4], s[4];
      const int iT = idx0 + e, iS = Q + idx0 + e;
      if (vec && iT + 3 < P.n_pix) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(rowp) + iT));
        t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) t[j] = iT + j < P.n_pix ? load_pixel<float>(rowp, P.in_dtype, iT + j) : 0.f;
      }
      if (vec && iS + 3 < P.n_pix) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(rowp) + iS));
        s[0] = v.x; s[1] = v.y; s[2] = v.z; s[3] = v.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) s[j] = iS + j < P.n_pix ? load_pixel<float>(rowp, P.in_dtype, iS + j) : 0.f;
      }
      T[e / 2] = pack2(t[0], t[1]); T[e / 2 + 1] = pack2(t[2], t[3]);
      S[e / 2] = pack2(s[0], s[1]); S[e / 2 + 1] = pack2(s[2], s[3]);
    }
  } else if (row) {
    // vector j = 2l + s is the chunk [l][s*Q, s*Q + Q) of Psi_n; slot k = l
    const float* pt = Pin + slot * 2 * Q + idx0;
#pragma unroll
    for (int e = 0; e < EL; e += 4) {
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f), s = t;
      if (act && idx0 + e < LEN) {
        t = *reinterpret_cast<const float4*>(pt + e);
        s = *reinterpret_cast<const float4*>(pt + Q + e);
      }
      T[e / 2] = pack2(t.x, t.y); T[e / 2 + 1] = pack2(t.z, t.w);
      S[e / 2] = pack2(s.x, s.y); S[e / 2 + 1] = pack2(s.z, s.w);
    }
  } else {
    // vector c = column c of M_n; element e' = s*B + l lives at Psi_n[l][s*Q + c]
#pragma unroll
    for (int e = 0; e < EL; e += 2) {
      float2 v0 = make_float2(0.f, 0.f), v1 = v0;
      const int i0 = idx0 + e, i1 = i0 + 1;
      if (act && i0 < LEN) {
        v0 = *reinterpret_cast<const float2*>(Pin + (i0 & (B - 1)) * 2 * Q + (i0 >> st.logB) * Q + 2 * slot);
        v1 = *reinterpret_cast<const float2*>(Pin + (i1 & (B - 1)) * 2 * Q + (i1 >> st.logB) * Q + 2 * slot);
      }
      T[e / 2] = pack2(v0.x, v1.x);
      S[e / 2] = pack2(v0.y, v1.y);
    }
  }
  {
    // U-row elements jj = 4 sub .. 4 sub + 3 (row mode: NV <= 16 so 4 per lane suffice)
    const int j0 = sub * 4, pT = 2 * slot, pS = 2 * slot + 1;
    const bool on = row && act;
    UT[0] = pack2(on && j0 == pT ? 1.f : 0.f, on && j0 + 1 == pT ? 1.f : 0.f);
    UT[1] = pack2(on && j0 + 2 == pT ? 1.f : 0.f, on && j0 + 3 == pT ? 1.f : 0.f);
    US[0] = pack2(on && j0 == pS ? 1.f : 0.f, on && j0 + 1 == pS ? 1.f : 0.f);
    US[1] = pack2(on && j0 + 2 == pS ? 1.f : 0.f, on && j0 + 3 == pS ? 1.f : 0.f);
  }
  // ---- Frobenius norm -> zero threshold ----
  float nT = group_dot<E2>(T, T, LP), nS = group_dot<E2>(S, S, LP);
  float fro = (sub == 0) ? nT + nS : 0.f;
  fro = warp_sum(fro);
  const float thr2 = kFloor2 * fro;

  int sweeps = 0;
  if (fro > 0.f) {
    sweeps = jacobi_regs<E2>(T, S, UT, US, nT, nS, thr2, lane, logLP, st.M, row);
    nT = group_dot<E2>(T, T, LP);
    nS = group_dot<E2>(S, S, LP);
  }
  // ---- stable descending rank among the NV vectors ----
  int rT = 0, rS = 0;
  {
    const int pT = 2 * slot, pS = 2 * slot + 1;
#pragma unroll 2
    for (int j = 0; j < NV; ++j) {
      const float v = __shfl_sync(FULL, (j & 1) ? nS : nT, (j >> 1) << logLP);
      rT += (v > nT) || (v == nT && j < pT);
      rS += (v > nS) || (v == nS && j < pS);
    }
  }
  const bool kT = act && nT > thr2 && rT < r_rt;
  const bool kS = act && nS > thr2 && rS < r_rt;
  if (sv) {
    if (sub == 0 && act) {
      if (rT < P.sv_width) sv[rT] = nT > thr2 ? sqrtf(nT) : 0.f;
      if (rS < P.sv_width) sv[rS] = nS > thr2 ? sqrtf(nS) : 0.f;
    }
    for (int k = NV + lane; k < P.sv_width; k += 32) sv[k] = 0.f;
  }
  if (P.sweeps_out && lane == 0) P.sweeps_out[(size_t)img * P.L + n] = sweeps;

  if (row) {
    // A_n[(s,l), rank] = U-row element jj = 2l + s
    if (act) {
      float ut[4], us[4];
      unpack2(UT[0], ut[0], ut[1]); unpack2(UT[1], ut[2], ut[3]);
      unpack2(US[0], us[0], us[1]); unpack2(US[1], us[2], us[3]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int jj = sub * 4 + e;
        const int l = jj >> 1, s = jj & 1;
        if (jj < NV && l < b_rt) {
          if (rT < r_rt) A[(s * b_rt + l) * r_rt + rT] = kT ? ut[e] : 0.f;
          if (rS < r_rt) A[(s * b_rt + l) * r_rt + rS] = kS ? us[e] : 0.f;
        }
      }
      // next Psi: row `rank` <- vector (zero when dropped)
      if (idx0 < LEN) {
#pragma unroll
        for (int e = 0; e < EL; e += 4) {
          float4 t, s;
          unpack2(T[e / 2], t.x, t.y); unpack2(T[e / 2 + 1], t.z, t.w);
          unpack2(S[e / 2], s.x, s.y); unpack2(S[e / 2 + 1], s.z, s.w);
          if (!kT) t = make_float4(0.f, 0.f, 0.f, 0.f);
          if (!kS) s = make_float4(0.f, 0.f, 0.f, 0.f);
          if (rT < R) *reinterpret_cast<float4*>(Pout + rT * Q + idx0 + e) = t;
          if (rS < R) *reinterpret_cast<float4*>(Pout + rS * Q + idx0 + e) = s;
        }
      }
    }
    __syncwarp();
    return Pout;
  }
  // ---- column mode: normalise -> U; stage U as [e'][R] in Pout; A_n to global ----
  {
    const float iT = kT ? rsqrtf(nT) : 0.f;
    const float iS = kS ? rsqrtf(nS) : 0.f;
    if (act) {
#pragma unroll
      for (int e = 0; e < EL; ++e) {
        const int idx = idx0 + e;
        float tv, sv_;
        {
          float lo, hi;
          unpack2(T[e / 2], lo, hi);
          tv = (e & 1) ? hi : lo;
          unpack2(S[e / 2], lo, hi);
          sv_ = (e & 1) ? hi : lo;
        }
        if (idx < LEN) {
          const int s = idx >> st.logB, l = idx & (B - 1);
          const float uT = tv * iT, uS = sv_ * iS;
          if (rT < R) Pout[idx * R + rT] = uT;
          if (rS < R) Pout[idx * R + rS] = uS;
          if (l < b_rt) {
            if (rT < r_rt) A[(s * b_rt + l) * r_rt + rT] = uT;
            if (rS < r_rt) A[(s * b_rt + l) * r_rt + rS] = uS;
          }
        }
      }
    }
    __syncwarp();
    if (n + 1 < P.L) {
      switch (st.rpl) {
        case 32: utm_product<32>(Pin, Pout, LEN, st.logB, st.logQ, R, lane); break;
        case 16: utm_product<16>(Pin, Pout, LEN, st.logB, st.logQ, R, lane); break;
        case 8: utm_product<8>(Pin, Pout, LEN, st.logB, st.logQ, R, lane); break;
        case 4: utm_product<4>(Pin, Pout, LEN, st.logB, st.logQ, R, lane); break;
        case 2: utm_product<2>(Pin, Pout, LEN, st.logB, st.logQ, R, lane); break;
        default: utm_product<1>(Pin, Pout, LEN, st.logB, st.logQ, R, lane); break;
      }
    }
    return Pin;
  }
}

__global__ void __launch_bounds__(128, 4)
encode_fast_kernel(const __grid_constant__ EncodeParams P, const __grid_constant__ FastPlan plan) {
  extern __shared__ __align__(16) float smem_f[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  float* b0 = smem_f + (size_t)warp * 2048;
  float* b1 = b0 + 1024;
  for (int64_t img = (int64_t)blockIdx.x * wpb + warp; img < P.N; img += (int64_t)gridDim.x * wpb) {
    float* mps = reinterpret_cast<float*>(P.mps_out) + (size_t)img * P.mps_stride;
    float* cur = b0;
#pragma unroll 1
    for (int n = 0; n < OC_MAX_L; ++n) {
      const FastStep& st = plan.st[n];
      float* other = cur == b0 ? b1 : b0;
      if (st.NV == 1) {
        // last step: M (2B x 1) -> A = M / |M|; the 1x1 S.V is dropped (unit norm)
        const int B = 1 << st.logB, LEN = 2 * B;
        const int b_rt = P.bonds[n];
        const int s = lane >> st.logB, l = lane & (B - 1);
        const float v = lane < LEN ? cur[l * 2 + s] : 0.f;
        const float nn = warp_sum(v * v);
        const float inv = nn > 0.f ? rsqrtf(nn) : 0.f;
        if (lane < LEN && l < b_rt) mps[P.site_off[n] + (s * b_rt + l)] = v * inv;
        if (P.svals_out) {
          float* sv = reinterpret_cast<float*>(P.svals_out) + ((size_t)img * P.L + n) * P.sv_width;
          for (int k = lane; k < P.sv_width; k += 32) sv[k] = k == 0 ? sqrtf(nn) : 0.f;
        }
        if (P.sweeps_out && lane == 0) P.sweeps_out[(size_t)img * P.L + n] = 0;
      } else if (st.el16) {
        cur = run_step<16>(P, st, n, img, cur, other, mps, lane);
      } else {
        cur = run_step<4>(P, st, n, img, cur, other, mps, lane);
      }
    }
    __syncwarp();
  }
}

inline int ilog2(int x) {
  int l = 0;
  while ((1 << l) < x) ++l;
  return l;
}

inline bool encode_fast_supported(const EncodeParams& P) { return P.L == 10; }

// Build the per-step table for L = 10 and bond cap D (padded to a power of two).
inline FastPlan encode_fast_plan(int D) {
  FastPlan plan;
  int DP = 1;
  while (DP < D) DP <<= 1;
  if (DP > 32) DP = 32;
  auto bond = [&](int n) {
    int b = 1 << (n < 10 - n ? n : 10 - n);
    return b < DP ? b : DP;
  };
  for (int n = 0; n < 10; ++n) {
    FastStep& s = plan.st[n];
    const int B = bond(n), Q = 1 << (9 - n), Pp = 2 * B;
    s.logB = ilog2(B);
    s.logQ = ilog2(Q);
    s.R = bond(n + 1);
    for (int attempt = 0; attempt < 2; ++attempt) {
      // rows are the vectors when M_n is wide; second attempt falls back to
      // columns if a lane cannot hold its share of the U row (4 elements)
      s.row = (Pp < Q) && attempt == 0;
      s.NV = s.row ? Pp : Q;
      s.LEN = s.row ? Q : Pp;
      s.M = s.NV / 2;
      // EL = 16 when every slot gets >= 1 lane and at least half the warp is busy
      int el = 16;
      if (s.LEN < 16 || (s.LEN / 16) * s.M < 16) el = 4;
      int lp = s.LEN / el;
      if (lp < 1) lp = 1;
      while (lp * s.M > 32) {  // more slots than lanes: longer chunks per lane
        if (el == 4) { el = 16; lp = s.LEN / 16 > 0 ? s.LEN / 16 : 1; } else lp >>= 1;
      }
      s.el16 = el == 16;
      s.logLP = ilog2(lp);
      if (!s.row || 4 * lp >= s.NV) break;
    }
    const int ng = Q >= 32 ? 1 : 32 / Q;
    s.rpl = s.R >= ng ? s.R / ng : 1;
  }
  return plan;
}

inline int encode_fast_launch(const EncodeParams& P, int sms, cudaStream_t st) {
  const int warps = 4;
  const size_t smem = (size_t)warps * 2048 * sizeof(float);
  const int64_t need = (P.N + warps - 1) / warps;
  const int grid = (int)(need < (int64_t)sms * 16 ? need : (int64_t)sms * 16);
  const FastPlan plan = encode_fast_plan(P.D);
  encode_fast_kernel<<<grid, warps * 32, smem, st>>>(P, plan);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

}  // namespace oc
