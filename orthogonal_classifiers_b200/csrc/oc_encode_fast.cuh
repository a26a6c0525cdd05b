// Specialised encoder for L = 10 (512 < n_pix <= 1024), fp32, one image per
// warp; the vectors being orthogonalised live in REGISTERS.
//
// Step n works on M_n (P = 2B rows, Q cols), B = b_n, Q = 2^(9-n).  Its NV
// vectors (rows when M_n is wide, columns otherwise; length LEN) are spread
// over the warp as "slots" of LP lanes; a slot owns the two vectors at line
// positions (2k, 2k+1) — register sets T and S, each lane holding EL
// consecutive elements of both.  One-sided (Hestenes) Jacobi, odd-even ordering:
//   round A: rotate (T_k, S_k)        — no communication
//   round B: rotate (S_k, T_{k+1})    — T moves one slot down and back
// every rotation also swaps the two line positions, so after NV rounds every
// pair has met once (one sweep).  Squared norms travel with the vectors and are
// updated by the rotation (refreshed from the data every sweep and whenever an
// update cancels), so a round costs ONE dot product.  Element-wise arithmetic
// is packed fma.rn.f32x2 (SASS FFMA2); rotation parameters are branch-free.
//
// STEP-MAJOR execution: the chain is cut into five kernels
//   head (steps 0-2) | step 3 | step 4 | step 5 | tail (steps 6-9)
// each launched over all images, Psi_n passing through a global workspace
// (4 KB per image, L2-resident for the chunk sizes used).  Every warp on the
// chip then runs the same few KB of fully static code.  The fused one-kernel
// versions stalled on instruction fetch (profiles/encode_r1.md): 16 warps per
// SM in 10 different steps do not fit the instruction cache.
#pragma once
#include "oc_generic.cuh"

namespace oc {

typedef unsigned long long u64;

constexpr float kTol2 = 1e-12f;    // rotate if (a.b)^2 > (1e-6)^2 |a|^2 |b|^2
#ifndef OC_QSTOP2
#define OC_QSTOP2 1e-7f
#endif
constexpr float kQStop2 = OC_QSTOP2;  // a sweep whose largest rotation had cos^2 below this is the last
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
__device__ __forceinline__ u64 shfl2(u64 v, int src) {
  float lo, hi;
  unpack2(v, lo, hi);
  lo = __shfl_sync(FULL, lo, src);
  hi = __shfl_sync(FULL, hi, src);
  return pack2(lo, hi);
}
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rsqrt_approx(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__host__ __device__ constexpr int cmin(int a, int b) { return a < b ? a : b; }
__host__ __device__ constexpr int cmax(int a, int b) { return a > b ? a : b; }
__host__ __device__ constexpr int clog2(int x) { return x <= 1 ? 0 : 1 + clog2(x / 2); }

// static description of step n of the chain with bond cap DPAD (power of two)
template <int DPAD, int N>
struct Step {
  static constexpr int B = cmin(cmin(1 << N, 1 << (10 - N)), DPAD);          // rows of Psi_n
  static constexpr int Q = 1 << (9 - N);                                       // half its columns
  static constexpr int R = cmin(cmin(1 << (N + 1), 1 << (9 - N)), DPAD);      // rows of Psi_{n+1}
  static constexpr int P = 2 * B;
  static constexpr bool ROW = P < Q;
  static constexpr int NV = ROW ? P : Q;
  static constexpr int LEN = ROW ? Q : P;
  static constexpr int M = NV / 2;                                             // active slots
  // lanes per slot: all 32 lanes when the vectors are long enough for >= 4
  // elements per lane, else 4 elements per lane on fewer lanes
  static constexpr int LP0 = NV >= 2 ? 64 / NV : 32;
  static constexpr int EL = (LEN / LP0 >= 4) ? LEN / LP0 : cmin(4, LEN);
  static constexpr int LP = LEN / EL;
  static constexpr int LOGLP = clog2(LP);
  static constexpr int E2 = EL / 2;
  static constexpr int UEL = ROW ? cmax(2, NV / LP) : 2;                      // U-row elements per lane (even)
  static constexpr int U2 = UEL / 2;
  static_assert(NV == 1 || (M * LP <= 32 && LP * EL == LEN && EL % 2 == 0), "bad lane mapping");
  static_assert(!ROW || UEL * LP >= NV, "U row does not fit");
};

#ifdef OC_DEBUG_COUNTERS
static __device__ int g_dbg_lost_dummy;
#define g_dbg_lost (*dbg_lost_ptr())
__device__ __forceinline__ int* dbg_lost_ptr() { extern __shared__ __align__(16) float smem_dbg[]; return reinterpret_cast<int*>(smem_dbg + 4 * 2048) + (threadIdx.x >> 5); }
#endif
template <int E2, int LOGLP>
__device__ __forceinline__ float group_dot(const u64 (&a)[E2], const u64 (&b)[E2]) {
  // independent partial sums keep the FFMA2 dependency chains short
  constexpr int NA = E2 >= 4 ? 4 : E2;
  u64 acc[NA];
#pragma unroll
  for (int e = 0; e < NA; ++e) acc[e] = mul2(a[e], b[e]);
#pragma unroll
  for (int e = NA; e < E2; ++e) acc[e % NA] = fma2(a[e], b[e], acc[e % NA]);
  float v = 0.f;
#pragma unroll
  for (int e = 0; e < NA; ++e) {
    float lo, hi;
    unpack2(acc[e], lo, hi);
    v += lo + hi;
  }
#pragma unroll
  for (int o = 0; o < LOGLP; ++o) v += __shfl_xor_sync(FULL, v, 1 << o);
  return v;
}

// rotate (a, b) with cached squared norms (na, nb); afterwards a/na hold the
// vector of the FIRST line position (b') and b/nb the second (a').
template <int E2, int U2, int LOGLP, bool ROW>
__device__ __forceinline__ void rotate_pair(u64 (&a)[E2], u64 (&b)[E2], u64 (&ua)[U2], u64 (&ub)[U2],
                                            float& na, float& nb, float thr2, bool idle, bool& again) {
  const float ab = group_dot<E2, LOGLP>(a, b);
  const float aa = na, bb = nb;
  const float d = aa * bb;
  const float ab2 = ab * ab;
  const bool rot = (ab2 > kTol2 * d) && (aa > thr2) && (bb > thr2) && !idle;
  again = again || (rot && ab2 > kQStop2 * d);
  // Branch-free rotation parameters with only two dependent MUFUs.  With
  // delta = bb - aa, gamma = 2 ab, w = sqrt(delta^2 + gamma^2), den = |delta| + w:
  //   tan = sgn(delta) gamma / den,  cos = den q,  sin = sgn(delta) gamma q,
  //   q = (den^2 + gamma^2)^(-1/2)   (one Newton step -> c^2 + s^2 = 1 to ~1 ulp,
  // whatever the error of the approximate sqrt).  |theta| <= pi/4.
  const float delta = bb - aa;
  const float gamma = rot ? ab + ab : 0.f;
  const float w = sqrt_approx(fmaf(delta, delta, gamma * gamma));
  const float den = fabsf(delta) + w;
  const float sg = copysignf(gamma, delta * gamma);  // sgn(delta) * gamma  (sgn(0) = +)
  const float X = fmaf(den, den, gamma * gamma);
  float q = rsqrt_approx(X);
  q = q * fmaf(-0.5f * X, q * q, 1.5f);
  const float t = rot ? sg * rcp_approx(den) : 0.f;
  // idle: both vectors stay where they are (the second picks up a sign)
  const float c = idle ? 0.f : (rot ? den * q : 1.f);
  const float s = idle ? 1.f : (rot ? sg * q : 0.f);
  // |a'|^2 = aa - t ab, |b'|^2 = bb + t ab; positions swap
  float n_first = idle ? aa : fmaf(t, ab, bb);
  float n_second = idle ? bb : fmaf(-t, ab, aa);
  const bool lost = rot && (n_second < kCancel * aa || n_first < kCancel * bb);
  const u64 c2 = pack2(c, c), s2 = pack2(s, s), ns2 = pack2(-s, -s);
#pragma unroll
  for (int e = 0; e < E2; ++e) {
    const u64 x = a[e], y = b[e];
    a[e] = fma2(s2, x, mul2(c2, y));
    b[e] = fma2(c2, x, mul2(ns2, y));
  }
  if (ROW) {
#pragma unroll
    for (int e = 0; e < U2; ++e) {
      const u64 x = ua[e], y = ub[e];
      ua[e] = fma2(s2, x, mul2(c2, y));
      ub[e] = fma2(c2, x, mul2(ns2, y));
    }
  }
  if (__any_sync(FULL, lost)) {  // warp-uniform, rare after the first sweeps
#ifdef OC_DEBUG_COUNTERS
    ++g_dbg_lost;
#endif
    n_first = group_dot<E2, LOGLP>(a, a);
    n_second = group_dot<E2, LOGLP>(b, b);
  }
  na = n_first;
  nb = n_second;
}

// `mlive` (warp-uniform, 1..M): only slots [0, mlive) hold non-zero vectors
// (rows of Psi_n dropped by the previous step are exactly zero and sit at the
// end), so the odd-even line is cut to 2*mlive positions: a sweep is 2*mlive
// rounds instead of 2*M.
template <int E2, int U2, int LOGLP, int M, bool ROW>
__device__ __forceinline__ int jacobi_regs(u64 (&T)[E2], u64 (&S)[E2], u64 (&UT)[U2], u64 (&US)[U2],
                                           float& nT, float& nS, float thr2, int lane, int mlive) {
  constexpr int LP = 1 << LOGLP;
  const int slot = lane >> LOGLP;
  const int src_dn = (lane + LP) & 31;
  const int src_up = (lane - LP) & 31;
  const bool idleB = slot >= mlive - 1;
  int sweeps = 0;
#pragma unroll 1
  while (sweeps < kMaxSweeps) {
    bool again = false;
    nT = group_dot<E2, LOGLP>(T, T);
    nS = group_dot<E2, LOGLP>(S, S);
#pragma unroll 2
    for (int rr = 0; rr < mlive; ++rr) {
      rotate_pair<E2, U2, LOGLP, ROW>(T, S, UT, US, nT, nS, thr2, false, again);
      if (M > 1) {
        u64 R[E2], UR[U2];
#pragma unroll
        for (int e = 0; e < E2; ++e) R[e] = shfl2(T[e], src_dn);
        float nR = __shfl_sync(FULL, nT, src_dn);
#pragma unroll
        for (int e = 0; e < U2; ++e) UR[e] = ROW ? shfl2(UT[e], src_dn) : 0ull;
        rotate_pair<E2, U2, LOGLP, ROW>(S, R, US, UR, nS, nR, thr2, idleB, again);
#pragma unroll
        for (int e = 0; e < E2; ++e) T[e] = shfl2(R[e], src_up);
        nT = __shfl_sync(FULL, nR, src_up);
        if (ROW) {
#pragma unroll
          for (int e = 0; e < U2; ++e) UT[e] = shfl2(UR[e], src_up);
        }
      }
    }
    ++sweeps;
    // (two vectors: one rotation orthogonalises them in exact arithmetic only; with graded norms the
    //  fp32 rotation leaves cos ~ 1e-4, so the usual criterion decides here as well)
    if (!__any_sync(FULL, again)) break;
  }
  return sweeps;
}

struct StepIO {
  float* A;        // site tensor A_n of this image (global)
  float* sv;       // svals row of this step (global) or nullptr
  int32_t* sweeps; // sweeps slot or nullptr
  int b_rt, r_rt;  // run-time bonds of the output layout (b_n(D), b_{n+1}(D))
  int sv_width;
};

// One step.  Psi_n is in shared memory `Pin` laid out [B][2Q] (or, for N == 0,
// the image in global memory).  Psi_{n+1} [R][Q] is left in shared memory; the
// function returns the buffer holding it (Pin in column mode, Pout in row mode).
template <int DPAD, int N>
__device__ __forceinline__ float* jstep(const EncodeParams& P, int64_t img, const StepIO& io,
                                        float* Pin, float* Pout, int lane) {
  using C = Step<DPAD, N>;
  constexpr int B = C::B, Q = C::Q, R = C::R, NV = C::NV, LEN = C::LEN, LP = C::LP, EL = C::EL;
  constexpr int E2 = C::E2, U2 = C::U2, UEL = C::UEL, LOGLP = C::LOGLP, M = C::M;
  constexpr bool ROW = C::ROW;
  const int b_rt = io.b_rt, r_rt = io.r_rt;
  float* A = io.A;

  if constexpr (NV == 1) {
    // last step: M (2B x 1) -> A = M / |M|; the 1x1 S.V is dropped (unit norm)
    const int s = lane / B, l = lane % B;
    const float v = lane < 2 * B ? Pin[l * 2 + s] : 0.f;
    const float nn = warp_sum(v * v);
    const float inv = nn > 0.f ? rsqrtf(nn) : 0.f;
    if (lane < 2 * B && l < b_rt) A[s * b_rt + l] = v * inv;
    if (io.sv) {
      for (int k = lane; k < io.sv_width; k += 32) io.sv[k] = k == 0 ? sqrtf(nn) : 0.f;
    }
    if (io.sweeps && lane == 0) *io.sweeps = 0;
    return Pin;
  } else {
    const int slot = lane >> LOGLP, sub = lane & (LP - 1);
    const bool act = slot < M;
    const int idx0 = sub * EL;
    u64 T[E2], S[E2], UT[U2], US[U2];
    // ---- load the two vectors of this slot ----
    if constexpr (N == 0) {
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
    } else if constexpr (ROW) {
      // vector j = 2l + s is the chunk [l][s*Q, s*Q + Q) of Psi_n; slot k = l
      const float* pt = Pin + slot * 2 * Q + idx0;
#pragma unroll
      for (int e = 0; e < EL; e += 4) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f), s = t;
        if (act) {
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
        if (act) {
          v0 = *reinterpret_cast<const float2*>(Pin + (i0 % B) * 2 * Q + (i0 / B) * Q + 2 * slot);
          v1 = *reinterpret_cast<const float2*>(Pin + (i1 % B) * 2 * Q + (i1 / B) * Q + 2 * slot);
        }
        T[e / 2] = pack2(v0.x, v1.x);
        S[e / 2] = pack2(v0.y, v1.y);
      }
    }
#pragma unroll
    for (int e = 0; e < U2; ++e) {
      const int j0 = sub * UEL + 2 * e, pT = 2 * slot, pS = 2 * slot + 1;
      const bool on = ROW && act;
      UT[e] = pack2(on && j0 == pT ? 1.f : 0.f, on && j0 + 1 == pT ? 1.f : 0.f);
      US[e] = pack2(on && j0 == pS ? 1.f : 0.f, on && j0 + 1 == pS ? 1.f : 0.f);
    }
    // ---- Frobenius norm -> zero threshold ----
    float nT = group_dot<E2, LOGLP>(T, T), nS = group_dot<E2, LOGLP>(S, S);
    float fro = (sub == 0) ? nT + nS : 0.f;
    fro = warp_sum(fro);
    const float thr2 = kFloor2 * fro;

    int sweeps = 0;
#ifdef OC_DEBUG_COUNTERS
    int dbg_mlive = 0;
    if (lane == 0) g_dbg_lost = 0;
    __syncwarp();
#endif
    if (fro > 0.f) {
      // highest slot holding a non-zero vector (exact zeros: padding / dropped rows)
      const unsigned livemask = __ballot_sync(FULL, act && (nT + nS) > 0.f);
      const int mlive = M == 1 ? 1 : ((31 - __clz(livemask)) >> LOGLP) + 1;
#ifdef OC_DEBUG_COUNTERS
      dbg_mlive = mlive;
#endif
      sweeps = jacobi_regs<E2, U2, LOGLP, M, ROW>(T, S, UT, US, nT, nS, thr2, lane, mlive);
      nT = group_dot<E2, LOGLP>(T, T);
      nS = group_dot<E2, LOGLP>(S, S);
    }
    // ---- stable descending rank among the NV vectors ----
    int rT = 0, rS = 0;
    {
      const int pT = 2 * slot, pS = 2 * slot + 1;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const float v = __shfl_sync(FULL, (j & 1) ? nS : nT, (j >> 1) << LOGLP);
        rT += (v > nT) || (v == nT && j < pT);
        rS += (v > nS) || (v == nS && j < pS);
      }
    }
    const bool kT = act && nT > thr2 && rT < r_rt;
    const bool kS = act && nS > thr2 && rS < r_rt;
    if (io.sv) {
      if (sub == 0 && act) {
        if (rT < io.sv_width) io.sv[rT] = nT > thr2 ? sqrtf(nT) : 0.f;
        if (rS < io.sv_width) io.sv[rS] = nS > thr2 ? sqrtf(nS) : 0.f;
      }
      for (int k = NV + lane; k < io.sv_width; k += 32) io.sv[k] = 0.f;
    }
#ifdef OC_DEBUG_COUNTERS
    if (io.sweeps && lane == 0) { *io.sweeps = sweeps + 100 * dbg_mlive + 10000 * g_dbg_lost; g_dbg_lost = 0; }
#else
    if (io.sweeps && lane == 0) *io.sweeps = sweeps;
#endif

    if constexpr (ROW) {
      // A_n[(s,l), rank] = U-row element jj = 2l + s
      if (act) {
#pragma unroll
        for (int e = 0; e < UEL; ++e) {
          float lo, hi;
          unpack2(UT[e / 2], lo, hi);
          const float ut = (e & 1) ? hi : lo;
          unpack2(US[e / 2], lo, hi);
          const float us = (e & 1) ? hi : lo;
          const int jj = sub * UEL + e;
          const int l = jj >> 1, s = jj & 1;
          if (jj < NV && l < b_rt) {
            if (rT < r_rt) A[(s * b_rt + l) * r_rt + rT] = kT ? ut : 0.f;
            if (rS < r_rt) A[(s * b_rt + l) * r_rt + rS] = kS ? us : 0.f;
          }
        }
        // next Psi: row `rank` <- vector (zero when dropped)
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
      __syncwarp();
      return Pout;
    } else {
      // column mode: normalise -> U; stage U as [e'][R] in Pout; A_n to global
      const float iT = kT ? rsqrtf(nT) : 0.f;
      const float iS = kS ? rsqrtf(nS) : 0.f;
      if (act) {
#pragma unroll
        for (int e = 0; e < EL; ++e) {
          const int idx = idx0 + e;
          float lo, hi;
          unpack2(T[e / 2], lo, hi);
          const float uT = ((e & 1) ? hi : lo) * iT;
          unpack2(S[e / 2], lo, hi);
          const float uS = ((e & 1) ? hi : lo) * iS;
          const int s = idx / B, l = idx % B;
          if (rT < R) Pout[idx * R + rT] = uT;
          if (rS < R) Pout[idx * R + rS] = uS;
          if (l < b_rt) {
            if (rT < r_rt) A[(s * b_rt + l) * r_rt + rT] = uT;
            if (rS < r_rt) A[(s * b_rt + l) * r_rt + rS] = uS;
          }
        }
      }
      __syncwarp();
      if constexpr (N + 1 < 10) {
        // Psi_{n+1}[r][c] = sum_e U[e][r] M[e][c];  M[e = s*B + l][c] = Pin[l][s*Q + c]
        constexpr int NG = Q >= 32 ? 1 : 32 / Q;   // lane groups sharing a column
        constexpr int RPL = R >= NG ? R / NG : 1;   // rows per lane
        const int c = lane % Q, g = lane / Q;
        const bool active = g * RPL < R;
        float acc[RPL];
#pragma unroll
        for (int i = 0; i < RPL; ++i) acc[i] = 0.f;
        if (active) {
#pragma unroll 4
          for (int e = 0; e < LEN; ++e) {
            const float m = Pin[(e % B) * 2 * Q + (e / B) * Q + c];
            const float* u = Pout + e * R + g * RPL;
            if constexpr (RPL % 4 == 0) {
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
      return Pin;
    }
  }
}

// compile-time loop over steps N .. N1-1
template <int DPAD, int N, int N1>
__device__ __forceinline__ float* run_steps(const EncodeParams& P, int64_t img, float* mps, float* cur,
                                            float* b0, float* b1, int lane) {
  if constexpr (N < N1) {
    StepIO io;
    io.A = mps + P.site_off[N];
    io.sv = P.svals_out ? reinterpret_cast<float*>(P.svals_out) + ((size_t)img * P.L + N) * P.sv_width : nullptr;
    io.sweeps = P.sweeps_out ? P.sweeps_out + (size_t)img * P.L + N : nullptr;
    io.b_rt = P.bonds[N];
    io.r_rt = P.bonds[N + 1];
    io.sv_width = P.sv_width;
    float* next = jstep<DPAD, N>(P, img, io, cur, cur == b0 ? b1 : b0, lane);
    return run_steps<DPAD, N + 1, N1>(P, img, mps, next, b0, b1, lane);
  } else {
    return cur;
  }
}

// Kernel for steps [N0, N1).  ws_in: Psi_{N0} of every image ([B][2Q] floats,
// image stride kWsStride) unless N0 == 0; ws_out: Psi_{N1} unless N1 == 10.
constexpr int kWsStride = 1024;
// scratch behind the two Psi buffers for the pairing sort of the half-warp chain:
// keys (1 byte / image) | perm (int32: 4 image slots per warp of the packed step 4, <= chunk/2 + 32 warps)
// | 12 x 32 ints of bins (counts, starts, cursors, totals; own counts and guest records of the packed step 4)
constexpr int kSortBins = 32;
inline size_t encode_perm_ints(int64_t chunk) { return (size_t)(2 * chunk + 4 * kSortBins + 4); }
inline size_t encode_sort_bytes(int64_t chunk) {
  return (size_t)((chunk + 255) & ~(int64_t)255) + encode_perm_ints(chunk) * sizeof(int32_t) +
         (size_t)12 * kSortBins * sizeof(int32_t) + 256;
}

template <int DPAD, int N0, int N1>
__global__ void __launch_bounds__(128, 6) encode_steps_kernel(const __grid_constant__ EncodeParams P,
                                                           const float* __restrict__ ws_in,
                                                           float* __restrict__ ws_out) {
  extern __shared__ __align__(16) float smem_f[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  float* b0 = smem_f + (size_t)warp * 2048;
  float* b1 = b0 + 1024;
  constexpr int IN_ELEMS = N0 > 0 ? Step<DPAD, cmin(N0, 9)>::B * 2 * Step<DPAD, cmin(N0, 9)>::Q : 0;
  constexpr int OUT_ELEMS = N1 < 10 ? Step<DPAD, cmin(N1, 9)>::B * 2 * Step<DPAD, cmin(N1, 9)>::Q : 0;
  for (int64_t img = (int64_t)blockIdx.x * wpb + warp; img < P.N; img += (int64_t)gridDim.x * wpb) {
    float* mps = reinterpret_cast<float*>(P.mps_out) + (size_t)img * P.mps_stride;
    if constexpr (N0 > 0) {
      const float4* src = reinterpret_cast<const float4*>(ws_in + (size_t)img * kWsStride);
      float4* dst = reinterpret_cast<float4*>(b0);
#pragma unroll
      for (int i = lane; i < IN_ELEMS / 4; i += 32) dst[i] = __ldcs(src + i);
      __syncwarp();
    }
    float* cur = run_steps<DPAD, N0, N1>(P, img, mps, b0, b0, b1, lane);
    if constexpr (N1 < 10) {
      __syncwarp();
      const float4* src = reinterpret_cast<const float4*>(cur);
      float4* dst = reinterpret_cast<float4*>(ws_out + (size_t)img * kWsStride);
#pragma unroll
      for (int i = lane; i < OUT_ELEMS / 4; i += 32) dst[i] = src[i];
    }
    __syncwarp();
  }
}

// optional per-launch CUDA events (oc_set_option("enc_events", 1)): a ring of
// the last kEvRing oc_encode_batch calls, 6 events each (before / after the five
// launches of the first chunk), recorded on the launching stream
constexpr int kEvRing = 16;
struct EncodeEvents {
  bool on = false;
  int ncalls = 0;
  cudaEvent_t ev[kEvRing][6] = {};
  void mark(int i, cudaStream_t st) {
    if (!on) return;
    cudaEvent_t& e = ev[ncalls % kEvRing][i];
    if (!e) cudaEventCreate(&e);
    cudaEventRecord(e, st);
    if (i == 5) ++ncalls;
  }
};
inline EncodeEvents& encode_events() {
  static EncodeEvents e;
  return e;
}

inline bool encode_fast_supported(const EncodeParams& P) { return P.L == 10; }

inline int& encode_warps_per_cta() {
  static int w = 4;
  return w;
}

template <int DPAD, int N0, int N1>
inline cudaError_t launch_steps(const EncodeParams& P, const float* ws_in, float* ws_out, int sms,
                                cudaStream_t st) {
  const int warps = encode_warps_per_cta();
#ifdef OC_DEBUG_COUNTERS
  const size_t smem = (size_t)4 * 2048 * sizeof(float) + 64;
#else
  const size_t smem = (size_t)warps * 2048 * sizeof(float);
#endif
  // one image per warp, one CTA per 4 images: per-image cost varies with the
  // sweep count, so the hardware CTA scheduler does the load balancing
  const int64_t need = (P.N + warps - 1) / warps;
  const int grid = (int)need;
  (void)sms;
  encode_steps_kernel<DPAD, N0, N1><<<grid, warps * 32, smem, st>>>(P, ws_in, ws_out);
  return cudaGetLastError();
}

template <int DPAD>
inline cudaError_t launch_chain(const EncodeParams& P, float* ws0, float* ws1, int sms, cudaStream_t st) {
  cudaError_t e;
  EncodeEvents& ee = encode_events();
  ee.mark(0, st);
  if ((e = launch_steps<DPAD, 0, 3>(P, nullptr, ws0, sms, st)) != cudaSuccess) return e;
  ee.mark(1, st);
  if ((e = launch_steps<DPAD, 3, 4>(P, ws0, ws1, sms, st)) != cudaSuccess) return e;
  ee.mark(2, st);
  if ((e = launch_steps<DPAD, 4, 5>(P, ws1, ws0, sms, st)) != cudaSuccess) return e;
  ee.mark(3, st);
  if ((e = launch_steps<DPAD, 5, 6>(P, ws0, ws1, sms, st)) != cudaSuccess) return e;
  ee.mark(4, st);
  e = launch_steps<DPAD, 6, 10>(P, ws1, nullptr, sms, st);
  ee.mark(5, st);
  return e;
}

constexpr int64_t kEncodeChunk = 131072;  // images per pass: 2 x 512 MB of Psi workspace
constexpr int kEncodeLaunchesPerChunk = 5;

inline int& encode_use_hw() {
  static int v = 1;
  return v;
}
// opt-in (oc_set_option("enc_split2", 1)): halves of a chunk on two streams
inline int& encode_split2() {
  static int v = 0;
  return v;
}
struct SplitStreams {
  cudaStream_t aux = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};
namespace hw {
inline bool supported(const EncodeParams& P);
inline cudaError_t launch_chain(const EncodeParams& P, float* ws0, float* ws1, void* sort_ws, int64_t chunk,
                                int sms, cudaStream_t st);
}  // namespace hw

// ws: 2 * min(N, kEncodeChunk) * kWsStride floats of device scratch.
// split: the caller's (per device and stream) auxiliary stream for "enc_split2", nullable.
inline int encode_fast_launch(const EncodeParams& P0, float* ws, int sms, cudaStream_t st,
                              SplitStreams* split = nullptr) {
  int DP = 1;
  while (DP < P0.D) DP <<= 1;
  if (DP > 32) DP = 32;
  const size_t in_esz = P0.in_dtype == 0 ? 4 : (P0.in_dtype == 1 ? 8 : 1);
  const int64_t chunk = P0.N < kEncodeChunk ? P0.N : kEncodeChunk;
  for (int64_t lo = 0; lo < P0.N; lo += kEncodeChunk) {
    EncodeParams P = P0;
    P.N = P0.N - lo < kEncodeChunk ? P0.N - lo : kEncodeChunk;
    P.images = reinterpret_cast<const char*>(P0.images) + (size_t)lo * P0.ld * in_esz;
    P.mps_out = reinterpret_cast<float*>(P0.mps_out) + (size_t)lo * P0.mps_stride;
    if (P0.svals_out) P.svals_out = reinterpret_cast<float*>(P0.svals_out) + (size_t)lo * P0.L * P0.sv_width;
    if (P0.sweeps_out) P.sweeps_out = P0.sweeps_out + (size_t)lo * P0.L;
    float* ws0 = ws;
    float* ws1 = ws + (size_t)chunk * kWsStride;
    cudaError_t e;
    const bool ev_on = encode_events().on;
    if (lo > 0) encode_events().on = false;  // events cover the first chunk of the call
    struct Restore {
      bool v;
      ~Restore() { encode_events().on = v; }
    } restore{ev_on};
    if (encode_use_hw() && hw::supported(P)) {
      char* sort_ws = reinterpret_cast<char*>(ws + (size_t)2 * chunk * kWsStride);
      if (split && encode_split2() && !ev_on && P.N >= 32768) {
        // two halves of the chunk on two streams: the CTAs of one half fill the SMs that the
        // last wave of the other half's kernel leaves idle (a Jacobi CTA lives 100-200 us)
        SplitStreams& ss = *split;
        if (!ss.aux) {
          if ((e = cudaStreamCreateWithFlags(&ss.aux, cudaStreamNonBlocking)) != cudaSuccess) return (int)e;
          cudaEventCreateWithFlags(&ss.fork, cudaEventDisableTiming);
          cudaEventCreateWithFlags(&ss.join, cudaEventDisableTiming);
        }
        const int64_t nA = ((P.N / 2) + 63) & ~(int64_t)63;
        EncodeParams PA = P, PB = P;
        PA.N = nA;
        PB.N = P.N - nA;
        PB.images = reinterpret_cast<const char*>(P.images) + (size_t)nA * P0.ld * in_esz;
        PB.mps_out = reinterpret_cast<float*>(P.mps_out) + (size_t)nA * P0.mps_stride;
        if (P.svals_out) PB.svals_out = reinterpret_cast<float*>(P.svals_out) + (size_t)nA * P0.L * P0.sv_width;
        if (P.sweeps_out) PB.sweeps_out = P.sweeps_out + (size_t)nA * P0.L;
        cudaEventRecord(ss.fork, st);
        cudaStreamWaitEvent(ss.aux, ss.fork, 0);
        e = hw::launch_chain(PB, ws0 + (size_t)nA * kWsStride, ws1 + (size_t)nA * kWsStride,
                             sort_ws + ((encode_sort_bytes(chunk) + 255) & ~(size_t)255), chunk, sms, ss.aux);
        if (e != cudaSuccess) return (int)e;
        e = hw::launch_chain(PA, ws0, ws1, sort_ws, chunk, sms, st);
        if (e != cudaSuccess) return (int)e;
        cudaEventRecord(ss.join, ss.aux);
        cudaStreamWaitEvent(st, ss.join, 0);
        continue;
      }
      e = hw::launch_chain(P, ws0, ws1, sort_ws, chunk, sms, st);
      if (e != cudaSuccess) return (int)e;
      continue;
    }
    switch (DP) {
      case 32: e = launch_chain<32>(P, ws0, ws1, sms, st); break;
      case 16: e = launch_chain<16>(P, ws0, ws1, sms, st); break;
      case 8: e = launch_chain<8>(P, ws0, ws1, sms, st); break;
      case 4: e = launch_chain<4>(P, ws0, ws1, sms, st); break;
      default: e = launch_chain<2>(P, ws0, ws1, sms, st); break;  // D <= 2 (D = 1: all bonds 1)
    }
    if (e != cudaSuccess) return (int)e;
  }
  return 0;
}

}  // namespace oc
