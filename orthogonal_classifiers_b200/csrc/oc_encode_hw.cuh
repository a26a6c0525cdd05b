// Half-warp encoder (fp32, L = 10, 16 < D <= 32): steps 0..5 of the chain.
//
// One image per HALF-warp (two images per warp).  Every step up to the
// 64 x 16 one has NV <= 32 vectors of 1024/NV*2 ... elements, i.e. exactly
// 1024 numbers, so each of the 16 lanes always holds 32 elements of the two
// vectors (T, S) of its slot in registers:
//
//   step  mode  vectors x length   slots  lanes/slot
//    0    row      2 x 512           1      16
//    1    row      4 x 256           2       8
//    2    row      8 x 128           4       4
//    3    row     16 x  64           8       2
//    4    row     32 x  32          16       1      (rows 26.. are zero for 784 px)
//    5    col     16 x  64           8       2
//
// Compared with the one-image-per-warp kernels (oc_encode_fast.cuh) the
// rotation-parameter chain and the shuffle reductions are shared by twice as
// many elements, and
//   * step 4 rotates ROWS (only 2*rank(bond 4) <= 26 of the 32 are non-zero,
//     ~20 for MNIST-like images, so the odd-even line is that short), and
//   * no step accumulates U: the converged rows ARE W = S.V^T (next Psi), and
//     A_n = U = M W^T S^-2 is one small product at the end of the step.  An
//     error in column k of U grows like sigma_max/sigma_k but enters every
//     gauge-invariant quantity weighted by sigma_k, so amplitudes, singular
//     values and overlaps keep fp32 accuracy (tests/test_gpu_parity.py).
// Steps run as separate kernels (head = 0..2 | 3 | 4 | 5), Psi passing through
// the global workspace like the v3 chain; steps 6..9 reuse the v3 tail kernel.
#pragma once
#include "oc_encode_fast.cuh"

#ifndef OC_HW_LIFT
#define OC_HW_LIFT 0  // experiment: in-place lifting rotations (measured: no faster, see profiles/encode_r1.md)
#endif

namespace oc {
namespace hw {

__device__ __forceinline__ u64 add2(u64 a, u64 b) {
  u64 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float hsum2(u64 v) {
  float lo, hi;
  unpack2(v, lo, hi);
  return lo + hi;
}
// 64-bit shared-memory accesses the compiler must not fuse into 128-bit ones
// (register pairs of the vector arrays are not quad-aligned: fusing costs MOVs)
__device__ __forceinline__ void sts64(unsigned addr, u64 v) {
  asm volatile("st.shared.b64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ u64 lds64(unsigned addr) {
  u64 v;
  asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <bool G>
__device__ __forceinline__ float4 ld4(const float* p) {
  if constexpr (G) return ldg4(p);
  else return lds4(p);
}

// dot product of two 32-element register vectors, summed over the LP lanes of a slot
// (NE < 16: only the first NE packed pairs can be non-zero)
template <int LOGLP, int NE = 16>
__device__ __forceinline__ float dot16(const u64 (&a)[16], const u64 (&b)[16]) {
  u64 acc[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) acc[e] = mul2(a[e], b[e]);
#pragma unroll
  for (int e = 4; e < NE; ++e) acc[e & 3] = fma2(a[e], b[e], acc[e & 3]);
  float v = hsum2(add2(add2(acc[0], acc[1]), add2(acc[2], acc[3])));
#pragma unroll
  for (int o = 0; o < LOGLP; ++o) v += __shfl_xor_sync(FULL, v, 1 << o);
  return v;
}

// Rotate (a, b) with cached squared norms; afterwards a/na hold the vector of
// the FIRST line position and b/nb the second (positions swap).  Same
// branch-free parameter chain as oc_encode_fast.cuh::rotate_pair.  `hold`
// makes the call an exact no-op for this lane (no rotation, no swap): line
// ends in round B, and everything a finished / shorter half-warp still has to
// sit through while its partner image works, so that the result for an image
// never depends on which image shares its warp.
template <int LOGLP, int NE = 16>
__device__ __forceinline__ void rotate16(u64 (&a)[16], u64 (&b)[16], float& na, float& nb, float thr2,
                                         bool hold, bool& lostf, bool& again) {
  const float ab = dot16<LOGLP, NE>(a, b);
  const float aa = na, bb = nb;
  const float d = aa * bb;
  const float ab2 = ab * ab;
  const bool rot = (ab2 > kTol2 * d) && (aa > thr2) && (bb > thr2) && !hold;
  again = again || (rot && ab2 > kQStop2 * d);
  const float delta = bb - aa;
  const float gamma = rot ? ab + ab : 0.f;
  const float w = sqrt_approx(fmaf(delta, delta, gamma * gamma));
  const float den = fabsf(delta) + w;
  const float sg = copysignf(gamma, delta * gamma);
  const float X = fmaf(den, den, gamma * gamma);
  float q = rsqrt_approx(X);
  q = q * fmaf(-0.5f * X, q * q, 1.5f);
  const float t = rot ? sg * rcp_approx(den) : 0.f;
  // a' = s x + c y,  b' = c x + ns y   (rotation + swap: ns = -s; hold: s = ns = 1, c = 0)
  const float c = hold ? 0.f : (rot ? den * q : 1.f);
  const float s = hold ? 1.f : (rot ? sg * q : 0.f);
  const float ns = hold ? 1.f : -s;
  float n_first = hold ? aa : fmaf(t, ab, bb);
  float n_second = hold ? bb : fmaf(-t, ab, aa);
  // a norm update that cancelled is refreshed from the data at the end of the
  // round pair (jacobi16); until then the stale value only makes one rotation
  // less effective, never wrong (c and s stay consistent)
  lostf = lostf || (rot && (n_second < kCancel * aa || n_first < kCancel * bb));
  const u64 c2 = pack2(c, c), s2 = pack2(s, s), ns2 = pack2(ns, ns);
#pragma unroll
  for (int e = 0; e < NE; ++e) {
    const u64 x = a[e], y = b[e];
    a[e] = fma2(s2, x, mul2(c2, y));
    b[e] = fma2(c2, x, mul2(ns2, y));
  }
  na = n_first;
  nb = n_second;
}

// Fast (scaled) rotation: the true vectors are a = da * A, b = db * B with A, B
// the stored arrays.  With t = tan(theta) the rotated pair is
//   first'  = c (b + t a) = (c db) (B + beta A),    beta  = t da / db
//   second' = c (a - t b) = (c da) (A - alpha B),   alpha = t db / da
// so the cosine goes into the scale factors and the update is ONE packed FMA
// per output pair instead of two.  Only t sits on the critical path (two
// dependent MUFUs); c = (1 + t^2)^(-1/2) is refined by a Newton step from the t
// actually used, so the transformation is orthogonal to ~1 ulp whatever the
// error of the approximate sqrt / rcp.  `hold`: exact no-op for this lane.
template <int LOGLP>
__device__ __forceinline__ void frot16(u64 (&a)[16], u64 (&b)[16], float& na, float& nb, float& da, float& db,
                                       float thr2, bool hold, unsigned hmask, bool& again) {
  const float r_ab = da * rcp_approx(db);  // da / db, db / da: off the critical path
  const float r_ba = db * rcp_approx(da);
  const float dd = da * db;
  const float ab = dot16<LOGLP>(a, b) * dd;
  const float aa = na, bb = nb;
  const float d = aa * bb;
  const float ab2 = ab * ab;
  const bool rot = (ab2 > kTol2 * d) && (aa > thr2) && (bb > thr2) && !hold;
  again = again || (rot && ab2 > kQStop2 * d);
  const float delta = bb - aa;
  const float gamma = rot ? ab + ab : 0.f;
  const float w = sqrt_approx(fmaf(delta, delta, gamma * gamma));
  const float den = fabsf(delta) + w;
  const float sg = copysignf(gamma, delta * gamma);
  const float t = rot ? sg * rcp_approx(den) : 0.f;
  const float beta = t * r_ab, alpha = t * r_ba;
  const float X = fmaf(t, t, 1.f);
  float c = rsqrt_approx(X);
  c = c * fmaf(-0.5f * X, c * c, 1.5f);
  bool lost = false;
  if (!hold) {
    const u64 b2 = pack2(beta, beta), na2 = pack2(-alpha, -alpha);
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const u64 x = a[e], y = b[e];
      a[e] = fma2(b2, x, y);
      b[e] = fma2(na2, y, x);
    }
    const float n_first = fmaf(t, ab, bb), n_second = fmaf(-t, ab, aa);
    lost = rot && (n_second < kCancel * aa || n_first < kCancel * bb);
    na = n_first;
    nb = n_second;
    const float d_first = c * db, d_second = c * da;
    da = d_first;
    db = d_second;
  }
  const unsigned lostmask = __ballot_sync(FULL, lost);
  if (lostmask) {  // warp-uniform, rare after the first sweeps
    const float f1 = dot16<LOGLP>(a, a) * da * da, f2 = dot16<LOGLP>(b, b) * db * db;
    if (lost) {
      na = f1;
      nb = f2;
    }
  }
}

// One-sided Jacobi, odd-even ordering, on the 2*M vectors of each half-warp.
// mlive: live slots of THIS half (0 = nothing to do); mlive_max: warp maximum.
#ifndef OC_HW_FAST_ROT
#define OC_HW_FAST_ROT 0
#endif
#ifndef OC_HW_UNROLL
#define OC_HW_UNROLL 1
#endif
constexpr int kHwUnroll = OC_HW_UNROLL;
template <int LOGLP, int NE = 16>
__device__ __forceinline__ int jacobi16(u64 (&T)[16], u64 (&S)[16], float& nT, float& nS, float thr2,
                                        int lane, int mlive, int mlive_max) {
  constexpr int LP = 1 << LOGLP;
  constexpr int M = 16 >> LOGLP;
  const int l16 = lane & 15, hb = lane & 16;
  const int slot = l16 >> LOGLP;
  const int src_dn = hb | ((l16 + LP) & 15);
  const int src_up = hb | ((l16 - LP) & 15);
  const bool idleB = slot >= mlive - 1;
  const unsigned hmask = 0xffffu << hb;
  int sweeps = 0;
  bool done = mlive == 0;
#if OC_HW_FAST_ROT
  float dT = 1.f, dS = 1.f;
#endif
#pragma unroll 1
  for (int sw = 0; sw < kMaxSweeps; ++sw) {
    bool again = false;
    bool lostf = false;
    {
#if OC_HW_FAST_ROT
      // fold the scale factors back into the data before refreshing the norms
      const u64 t2 = pack2(dT, dT), s2 = pack2(dS, dS);
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        T[e] = mul2(T[e], t2);
        S[e] = mul2(S[e], s2);
      }
      dT = dS = 1.f;
#endif
      const float fT = dot16<LOGLP, NE>(T, T), fS = dot16<LOGLP, NE>(S, S);
      if (!done) {
        nT = fT;
        nS = fS;
      }
    }
#pragma unroll kHwUnroll
    for (int rr = 0; rr < mlive_max; ++rr) {
      const bool off = done || rr >= mlive;  // this half has nothing left to do in this round
#if OC_HW_FAST_ROT
      frot16<LOGLP>(T, S, nT, nS, dT, dS, thr2, off, hmask, again);
#else
      rotate16<LOGLP, NE>(T, S, nT, nS, thr2, off, lostf, again);
#endif
      if (M > 1) {
#pragma unroll
        for (int e = 0; e < NE; ++e) T[e] = shfl2(T[e], src_dn);
        float nR = __shfl_sync(FULL, nT, src_dn);
#if OC_HW_FAST_ROT
        float dR = __shfl_sync(FULL, dT, src_dn);
        frot16<LOGLP>(S, T, nS, nR, dS, dR, thr2, off || idleB, hmask, again);
        dT = __shfl_sync(FULL, dR, src_up);
#else
        rotate16<LOGLP, NE>(S, T, nS, nR, thr2, off || idleB, lostf, again);
#endif
#pragma unroll
        for (int e = 0; e < NE; ++e) T[e] = shfl2(T[e], src_up);
        nT = __shfl_sync(FULL, nR, src_up);
      }
#if !OC_HW_FAST_ROT
      {
        const unsigned lostmask = __ballot_sync(FULL, lostf);
        if (lostmask) {  // warp-uniform, rare after the first sweeps
          const float fT = dot16<LOGLP, NE>(T, T), fS = dot16<LOGLP, NE>(S, S);
          if (lostmask & hmask) {  // only the half that cancelled takes the refreshed norms
            nT = fT;
            nS = fS;
          }
          lostf = false;
        }
      }
#endif
    }
    if (!done) ++sweeps;
    const unsigned b = __ballot_sync(FULL, again);
    if (M == 1 || mlive <= 1 || !(b & hmask)) done = true;
    if (__all_sync(FULL, done)) break;
  }
#if OC_HW_FAST_ROT
  {
    const u64 t2 = pack2(dT, dT), s2 = pack2(dS, dS);
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      T[e] = mul2(T[e], t2);
      S[e] = mul2(S[e], s2);
    }
  }
#endif
  return sweeps;
}

// ---------------------------------------------------------------------------
// In-place ("lifting") form of the fast rotation.  With x at the first and y at
// the second position of the pair, t = tan(theta), c = cos(theta):
//     y <- y + beta x            beta  = t dx / dy          (scale dy <- c dy)
//     x <- x - alpha y_new       alpha = t c^2 dy / dx      (scale dx <- dx / c)
// leaves c (y + t x) in y's registers and c (x - t y) in x's, two packed FMAs
// per element pair, no temporaries and no register renaming; beta = alpha = 0
// is an exact no-op, so holding a lane costs nothing.  The vectors do NOT swap
// registers; the position swap of the odd-even ordering is done by the data
// movement in jacobi16_lift.
template <int LOGLP>
__device__ __forceinline__ void lift16(u64 (&x)[16], u64 (&y)[16], float& nx, float& ny, float& dx, float& dy,
                                       float thr2, bool hold, bool& lostf, bool& again) {
  const float r_xy = dx * rcp_approx(dy);  // off the critical path
  const float r_yx = dy * rcp_approx(dx);
  const float ab = dot16<LOGLP>(x, y) * (dx * dy);
  const float aa = nx, bb = ny;
  const float d = aa * bb;
  const float ab2 = ab * ab;
  const bool rot = (ab2 > kTol2 * d) && (aa > thr2) && (bb > thr2) && !hold;
  again = again || (rot && ab2 > kQStop2 * d);
  const float delta = bb - aa;
  const float gamma = rot ? ab + ab : 0.f;
  const float w = sqrt_approx(fmaf(delta, delta, gamma * gamma));
  const float den = fabsf(delta) + w;
  const float sg = copysignf(gamma, delta * gamma);
  const float t = rot ? sg * rcp_approx(den) : 0.f;
  const float X = fmaf(t, t, 1.f);
  float c = rsqrt_approx(X);
  c = c * fmaf(-0.5f * X, c * c, 1.5f);  // Newton: the transformation is orthogonal to ~1 ulp for the t used
  const float beta = rot ? t * r_xy : 0.f;
  const float alpha = rot ? t * r_yx * (c * c) : 0.f;
  const u64 b2 = pack2(beta, beta), na2 = pack2(-alpha, -alpha);
#pragma unroll
  for (int e = 0; e < 16; ++e) y[e] = fma2(b2, x[e], y[e]);
#pragma unroll
  for (int e = 0; e < 16; ++e) x[e] = fma2(na2, y[e], x[e]);
  const float n_y = fmaf(t, ab, bb), n_x = fmaf(-t, ab, aa);  // t = 0: unchanged
  lostf = lostf || (rot && (n_x < kCancel * aa || n_y < kCancel * bb));
  nx = n_x;
  ny = n_y;
  dy = rot ? c * dy : dy;
  dx = rot ? dx * (X * c) : dx;
}

// Odd-even ordering with the lifting rotation.  Register sets P, Q of slot k
// hold line positions (2k, 2k+1) between iterations.  One iteration:
//   A: rotate (P_k, Q_k)                    -> logically P = second, Q = first
//   Q moves one slot down (lane k gets first_{k+1})
//   B: rotate (P_k, Q_k) = (second_k, first_{k+1})
//   P moves one slot up (lane k+1 gets its new first position)
// The two vectors that sit out round B (first_0 and second_{m-1}) change sets at
// the line ends through a per-lane scratch row in shared memory (written and
// read by the same lane).  A half that is finished, or past its own line length
// in this sweep, neither rotates nor moves (shuffle source = itself), so its
// state never depends on the image sharing its warp.
// refl: refl_floats(LOGLP) floats of shared memory of this IMAGE (each lane uses its own rows).
template <int LOGLP>
__device__ __forceinline__ int jacobi16_lift(u64 (&P)[16], u64 (&Q)[16], float& nP, float& nQ, float thr2,
                                             int lane, int mlive, int mlive_max, float* refl) {
  constexpr int LP = 1 << LOGLP;
  constexpr int M = 16 >> LOGLP;
  const int l16 = lane & 15, hb = lane & 16;
  const int slot = l16 >> LOGLP;
  const int src_dn = hb | ((l16 + LP) & 15);
  const int src_up = hb | ((l16 - LP) & 15);
  const bool live = slot < mlive;
  const bool first = live && slot == 0, last = live && slot == mlive - 1;
  const bool idleB = slot >= mlive - 1;
  const unsigned hmask = 0xffffu << hb;
  int sweeps = 0;
  bool done = mlive == 0;
  float dP = 1.f, dQ = 1.f;
  const unsigned rq = (unsigned)__cvta_generic_to_shared(refl + (l16 & (LP - 1)) * 32);
  const unsigned rp = (unsigned)__cvta_generic_to_shared(refl + (LP + (l16 & (LP - 1))) * 32);
#pragma unroll 1
  for (int sw = 0; sw < kMaxSweeps; ++sw) {
    bool again = false;
    bool lostf = false;
    {
      // fold the scale factors back into the data, refresh the norms
      const u64 p2 = pack2(dP, dP), q2 = pack2(dQ, dQ);
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        P[e] = mul2(P[e], p2);
        Q[e] = mul2(Q[e], q2);
      }
      dP = dQ = 1.f;
      const float fP = dot16<LOGLP>(P, P), fQ = dot16<LOGLP>(Q, Q);
      if (!done && live) {
        nP = fP;
        nQ = fQ;
      }
    }
#pragma unroll 1
    for (int rr = 0; rr < mlive_max; ++rr) {
      const bool off = done || rr >= mlive;  // this half sits this round pair out
      lift16<LOGLP>(P, Q, nP, nQ, dP, dQ, thr2, off || !live, lostf, again);
      if (M > 1) {
        const bool mvf = first && !off, mvl = last && !off;
        const int sq = off ? lane : src_dn, sp = off ? lane : src_up;
        const float keepN = nQ, keepD = dQ;
        if (mvf) {
#pragma unroll
          for (int e = 0; e < 16; ++e) sts64(rq + 8 * e, Q[e]);
        }
#pragma unroll
        for (int e = 0; e < 16; ++e) Q[e] = shfl2(Q[e], sq);
        nQ = __shfl_sync(FULL, nQ, sq);
        dQ = __shfl_sync(FULL, dQ, sq);
        lift16<LOGLP>(P, Q, nP, nQ, dP, dQ, thr2, off || idleB, lostf, again);
        const float keepN2 = nP, keepD2 = dP;
        if (mvl) {
#pragma unroll
          for (int e = 0; e < 16; ++e) sts64(rp + 8 * e, P[e]);
        }
#pragma unroll
        for (int e = 0; e < 16; ++e) P[e] = shfl2(P[e], sp);
        nP = __shfl_sync(FULL, nP, sp);
        dP = __shfl_sync(FULL, dP, sp);
        if (mvf) {
#pragma unroll
          for (int e = 0; e < 16; ++e) P[e] = lds64(rq + 8 * e);
          nP = keepN;
          dP = keepD;
        }
        if (mvl) {
#pragma unroll
          for (int e = 0; e < 16; ++e) Q[e] = lds64(rp + 8 * e);
          nQ = keepN2;
          dQ = keepD2;
        }
      }
      {
        const unsigned lostmask = __ballot_sync(FULL, lostf);
        if (lostmask) {  // warp-uniform, rare after the first sweeps
          const float fP = dot16<LOGLP>(P, P) * dP * dP, fQ = dot16<LOGLP>(Q, Q) * dQ * dQ;
          if ((lostmask & hmask) && live && !off) {  // only the half that cancelled takes the refreshed norms
            nP = fP;
            nQ = fQ;
          }
          lostf = false;
        }
      }
    }
    if (!done) ++sweeps;
    const unsigned b = __ballot_sync(FULL, again);
    if (M == 1 || mlive <= 1 || !(b & hmask)) done = true;
    if (__all_sync(FULL, done)) break;
  }
  {
    // dead slots may have collected copies of the end vectors: clear them
    const u64 p2 = live ? pack2(dP, dP) : 0ull, q2 = live ? pack2(dQ, dQ) : 0ull;
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      P[e] = live ? mul2(P[e], p2) : 0ull;
      Q[e] = live ? mul2(Q[e], q2) : 0ull;
    }
  }
  return sweeps;
}

// One odd-even sweep of Gram-Schmidt projections on the columns of U held by
// the sub == 0 lane of every slot (B2 packed pairs each, unit norm up to the
// error being removed).  Of every pair that meets, the column of the SMALLER
// singular value (key) is orthogonalised against the other; nothing else
// changes, so the well-determined columns keep their accuracy.  Keys and rank
// tags travel with the columns (positions swap as in the Jacobi ordering).
template <int LOGLP, int B2>
__device__ __forceinline__ void polish_pair(u64 (&a)[B2], u64 (&b)[B2], float& ka, float& kb, int& ra, int& rb,
                                            bool hold) {
  u64 acc0 = mul2(a[0], b[0]), acc1 = 0ull;
  if (B2 > 1) acc1 = mul2(a[1], b[1]);
#pragma unroll
  for (int e = 2; e < B2; ++e) {
    if (e & 1) acc1 = fma2(a[e], b[e], acc1);
    else acc0 = fma2(a[e], b[e], acc0);
  }
  const float g = hsum2(add2(acc0, acc1));
  // a' = s x + c y,  b' = c x + ns y  (swap);  hold: identity
  const bool a_big = ka >= kb;
  const float c = hold ? 0.f : 1.f;
  const float s = hold ? 1.f : (a_big ? -g : 0.f);
  const float ns = hold ? 1.f : (a_big ? 0.f : -g);
  const u64 c2 = pack2(c, c), s2 = pack2(s, s), ns2 = pack2(ns, ns);
#pragma unroll
  for (int e = 0; e < B2; ++e) {
    const u64 x = a[e], y = b[e];
    a[e] = fma2(s2, x, mul2(c2, y));
    b[e] = fma2(c2, x, mul2(ns2, y));
  }
  if (!hold) {
    const float k = ka;
    ka = kb;
    kb = k;
    const int r = ra;
    ra = rb;
    rb = r;
  }
}

template <int LOGLP, int B2>
__device__ __forceinline__ void polish16(u64 (&T)[B2], u64 (&S)[B2], float& kT, float& kS, int& rT, int& rS,
                                         int lane, int mlive, int mlive_max) {
  constexpr int LP = 1 << LOGLP;
  constexpr int M = 16 >> LOGLP;
  const int l16 = lane & 15, hb = lane & 16;
  const int slot = l16 >> LOGLP;
  const int src_dn = hb | ((l16 + LP) & 15);
  const int src_up = hb | ((l16 - LP) & 15);
  const bool idleB = slot >= mlive - 1;
#pragma unroll 1
  for (int rr = 0; rr < mlive_max; ++rr) {
    const bool off = rr >= mlive;
    polish_pair<LOGLP, B2>(T, S, kT, kS, rT, rS, off);
    if (M > 1) {
#pragma unroll
      for (int e = 0; e < B2; ++e) T[e] = shfl2(T[e], src_dn);
      float kR = __shfl_sync(FULL, kT, src_dn);
      int rR = __shfl_sync(FULL, rT, src_dn);
      polish_pair<LOGLP, B2>(S, T, kS, kR, rS, rR, off || idleB);
#pragma unroll
      for (int e = 0; e < B2; ++e) T[e] = shfl2(T[e], src_up);
      kT = __shfl_sync(FULL, kR, src_up);
      rT = __shfl_sync(FULL, rR, src_up);
    }
  }
}

// stable descending rank of this lane's two vectors among the NV = 2*M of its half
template <int LOGLP>
__device__ __forceinline__ void rank16(float nT, float nS, int lane, int& rT, int& rS) {
  constexpr int NV = 2 * (16 >> LOGLP);
  const int hb = lane & 16, slot = (lane & 15) >> LOGLP;
  const int pT = 2 * slot, pS = pT + 1;
  rT = 0;
  rS = 0;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const float v = __shfl_sync(FULL, (j & 1) ? nS : nT, hb | ((j >> 1) << LOGLP));
    rT += (v > nT) || (v == nT && j < pT);
    rS += (v > nS) || (v == nS && j < pS);
  }
}

__device__ __forceinline__ float half_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

struct Out {
  float* A;         // site tensor A_n of this image
  float* sv;        // svals row of this step or nullptr
  int32_t* sweeps;  // sweeps slot or nullptr
  int b_rt, r_rt, sv_width;
};

__device__ __forceinline__ Out make_out(const EncodeParams& P, int64_t img, int n) {
  Out o;
  float* mps = reinterpret_cast<float*>(P.mps_out) + (size_t)img * P.mps_stride;
  o.A = mps + P.site_off[n];
  o.sv = P.svals_out ? reinterpret_cast<float*>(P.svals_out) + ((size_t)img * P.L + n) * P.sv_width : nullptr;
  o.sweeps = P.sweeps_out ? P.sweeps_out + (size_t)img * P.L + n : nullptr;
  o.b_rt = P.bonds[n];
  o.r_rt = P.bonds[n + 1];
  o.sv_width = P.sv_width;
  return o;
}

// floats of shared memory one image needs for the end-of-line scratch of jacobi16_lift
__host__ __device__ constexpr int refl_floats(int loglp) { return OC_HW_LIFT ? (64 << loglp) : 0; }
// floats of shared memory one image needs for the U columns of ROW step n
__host__ __device__ constexpr int ysm_floats(int n) { return (2 << n) * ((2 << n) + 2); }

// ROW step n (0..4): Psi_n [B][2Q] at psi_in (global if GSRC else shared),
// Psi_{n+1} [2B][Q] to psi_out (global if GDST else shared; may alias psi_in
// when shared: every read of psi_in happens before the first write).
// ysm: ysm_floats(N) floats of shared memory for this image.
// Everything that is only needed after the Jacobi loop (output pointers, key
// slot) is derived from P (constant bank) and the image index afterwards, so
// that it does not occupy registers across the loop.
template <int N, bool GSRC, bool GDST, bool KEYS = false>
__device__ __forceinline__ void row_step(bool valid, const EncodeParams& P, int64_t im, const float* psi_in,
                                         float* psi_out, float* ysm, float* refl, int lane,
                                         uint8_t* keys = nullptr) {
  constexpr int B = 1 << N, Q = 1 << (9 - N);
  constexpr int LOGLP = 4 - N, LP = 1 << LOGLP;
  constexpr int YSTR = 2 * B + 2;
  const int l16 = lane & 15, hb = lane & 16;
  const int slot = l16 >> LOGLP, sub = l16 & (LP - 1);
  u64 T[16], S[16];
  {
    const float* pt = psi_in + slot * 2 * Q + 4 * sub;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f), s = t;
      if (valid) {
        t = ld4<GSRC>(pt + 4 * LP * i);
        s = ld4<GSRC>(pt + Q + 4 * LP * i);
      }
      T[2 * i] = pack2(t.x, t.y); T[2 * i + 1] = pack2(t.z, t.w);
      S[2 * i] = pack2(s.x, s.y); S[2 * i + 1] = pack2(s.z, s.w);
    }
  }
  float nT = dot16<LOGLP>(T, T), nS = dot16<LOGLP>(S, S);
  const float fro = half_sum(sub == 0 ? nT + nS : 0.f);
  const float thr2 = kFloor2 * fro;
  // live slots: rows dropped by the previous step / padding are exactly zero and sit at the end
  const unsigned livemask = __ballot_sync(FULL, (nT + nS) > 0.f);
  const unsigned lm0 = livemask & 0xffffu, lm1 = livemask >> 16;
  const int ml0 = lm0 ? ((31 - __clz(lm0)) >> LOGLP) + 1 : 0;
  const int ml1 = lm1 ? ((31 - __clz(lm1)) >> LOGLP) + 1 : 0;
  const int mlive = hb ? ml1 : ml0;
  const int mlive_max = ml0 > ml1 ? ml0 : ml1;
  int sweeps = 0;
  if (mlive_max > 0) {
#if OC_HW_LIFT
    sweeps = jacobi16_lift<LOGLP>(T, S, nT, nS, thr2, lane, mlive, mlive_max, refl);
#else
    sweeps = jacobi16<LOGLP>(T, S, nT, nS, thr2, lane, mlive, mlive_max);
#endif
    nT = dot16<LOGLP>(T, T);
    nS = dot16<LOGLP>(S, S);
  }
  int rT, rS;
  rank16<LOGLP>(nT, nS, lane, rT, rS);
  const Out io = make_out(P, im, N);
  const int r_rt = io.r_rt, b_rt = io.b_rt;
  const bool kT = nT > thr2 && rT < r_rt;
  const bool kS = nS > thr2 && rS < r_rt;
  if (valid) {
    if (io.sv) {
      if (sub == 0) {
        if (rT < io.sv_width) io.sv[rT] = nT > thr2 ? sqrtf(nT) : 0.f;
        if (rS < io.sv_width) io.sv[rS] = nS > thr2 ? sqrtf(nS) : 0.f;
      }
      for (int k = 2 * B + l16; k < io.sv_width; k += 16) io.sv[k] = 0.f;
    }
    if (io.sweeps && l16 == 0) *io.sweeps = sweeps;
  }
  if constexpr (KEYS) {
    // number of kept vectors = non-zero rows of Psi_{n+1} = live slots of the next step
    const unsigned kt = __ballot_sync(FULL, kT && sub == 0), ks = __ballot_sync(FULL, kS && sub == 0);
    const int kept = __popc((kt >> hb) & 0xffffu) + __popc((ks >> hb) & 0xffffu);
    if (valid && l16 == 0) keys[im] = (uint8_t)kept;
  }
  // ---- U[(l,s), k] = (1/n_k) sum_c Psi_n[l][s*Q + c] W_k[c]  -> shared, column-major ----
  float* yT = ysm + slot * YSTR;
  float* yS = ysm + (B + slot) * YSTR;
  {
    const float iT = kT ? 1.f / nT : 0.f;
    const float iS = kS ? 1.f / nS : 0.f;
#pragma unroll 1
    for (int l = 0; l < B; ++l) {
      float aT[2] = {0.f, 0.f}, aS[2] = {0.f, 0.f};
      if (l < mlive_max) {  // warp-uniform; rows l >= mlive are zero
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const float* mr = psi_in + l * 2 * Q + s * Q + 4 * sub;
          u64 t0 = 0ull, t1 = 0ull, s0 = 0ull, s1 = 0ull;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) m = ld4<GSRC>(mr + 4 * LP * i);
            const u64 m0 = pack2(m.x, m.y), m1 = pack2(m.z, m.w);
            t0 = fma2(m0, T[2 * i], t0);
            t1 = fma2(m1, T[2 * i + 1], t1);
            s0 = fma2(m0, S[2 * i], s0);
            s1 = fma2(m1, S[2 * i + 1], s1);
          }
          float vt = hsum2(add2(t0, t1)), vs = hsum2(add2(s0, s1));
#pragma unroll
          for (int o = 0; o < LOGLP; ++o) {
            vt += __shfl_xor_sync(FULL, vt, 1 << o);
            vs += __shfl_xor_sync(FULL, vs, 1 << o);
          }
          aT[s] = vt * iT;
          aS[s] = vs * iS;
        }
      }
      if (sub == 0) {
        *reinterpret_cast<float2*>(yT + 2 * l) = make_float2(aT[0], aT[1]);
        *reinterpret_cast<float2*>(yS + 2 * l) = make_float2(aS[0], aS[1]);
      }
    }
  }
  if constexpr (!GDST) __syncwarp();  // all reads of psi_in done before it is overwritten
  // ---- Psi_{n+1}: row `rank` <- vector (zero when dropped) ----
  if (valid || !GDST) {
    float* pT = psi_out + rT * Q + 4 * sub;
    float* pS = psi_out + rS * Q + 4 * sub;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 t, s;
      unpack2(T[2 * i], t.x, t.y); unpack2(T[2 * i + 1], t.z, t.w);
      unpack2(S[2 * i], s.x, s.y); unpack2(S[2 * i + 1], s.z, s.w);
      if (!kT) t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!kS) s = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(pT + 4 * LP * i) = t;
      *reinterpret_cast<float4*>(pS + 4 * LP * i) = s;
    }
  }
  __syncwarp();
  // ---- polish U (see polish_pair), normalise, A_n[(s*b + l), rank] ----
  {
    u64 UT[B], US[B];
#pragma unroll
    for (int e = 0; e < B; ++e) {
      UT[e] = sub == 0 ? *reinterpret_cast<const u64*>(yT + 2 * e) : 0ull;
      US[e] = sub == 0 ? *reinterpret_cast<const u64*>(yS + 2 * e) : 0ull;
    }
    float keyT = kT ? nT : 0.f, keyS = kS ? nS : 0.f;
    int tagT = rT, tagS = rS;
    polish16<LOGLP, B>(UT, US, keyT, keyS, tagT, tagS, lane, mlive, mlive_max);
    u64 qT = 0ull, qS = 0ull;
#pragma unroll
    for (int e = 0; e < B; ++e) {
      qT = fma2(UT[e], UT[e], qT);
      qS = fma2(US[e], US[e], qS);
    }
    const float jT = keyT > 0.f ? rsqrtf(hsum2(qT)) : 0.f;
    const float jS = keyS > 0.f ? rsqrtf(hsum2(qS)) : 0.f;
    if (valid && sub == 0) {
#pragma unroll
      for (int l = 0; l < B; ++l) {
        float t0, t1, s0, s1;
        unpack2(UT[l], t0, t1);
        unpack2(US[l], s0, s1);
        if (l < b_rt) {
          if (tagT < r_rt) {
            io.A[l * r_rt + tagT] = t0 * jT;
            io.A[(b_rt + l) * r_rt + tagT] = t1 * jT;
          }
          if (tagS < r_rt) {
            io.A[l * r_rt + tagS] = s0 * jS;
            io.A[(b_rt + l) * r_rt + tagS] = s1 * jS;
          }
        }
      }
    }
  }
  __syncwarp();
}

#ifndef OC_HW_MINB
#define OC_HW_MINB 4
#endif
constexpr int kHwWarps = 4;  // warps per CTA (8 images)

// ---- head: steps 0..2, image -> Psi_3 (global workspace) ----
__global__ void __launch_bounds__(kHwWarps * 32, OC_HW_MINB) hw_head_kernel(const __grid_constant__ EncodeParams P,
                                                                   float* __restrict__ ws_out) {
  extern __shared__ __align__(16) float smem_f[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int l16 = lane & 15, h = lane >> 4;
  const int64_t img = ((int64_t)blockIdx.x * kHwWarps + warp) * 2 + h;
  const bool valid = img < P.N;
  float* buf = smem_f + (size_t)(warp * 2 + h) * 1024;
  float* ysm = smem_f + (size_t)kHwWarps * 2 * 1024 + (size_t)(warp * 2 + h) * ysm_floats(2);
  float* refl = smem_f + (size_t)kHwWarps * 2 * (1024 + ysm_floats(2)) + (size_t)(warp * 2 + h) * refl_floats(3);
  // stage the tail-padded image (tools.py:218-220: flat padding, pixel i -> amplitude i)
  {
    const char* src = reinterpret_cast<const char*>(P.images);
    const int esz = P.in_dtype == 0 ? 4 : (P.in_dtype == 1 ? 8 : 1);
    const void* rowp = src + (size_t)(valid ? img : 0) * P.ld * esz;
    const bool vec = P.in_dtype == 0 && ((P.ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    const bool vec8 = P.in_dtype == 2 && ((P.ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 3) == 0);
#pragma unroll 4
    for (int i = 4 * l16; i < 1024; i += 64) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid) {
        if (vec && i + 3 < P.n_pix) {
          v = __ldcs(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(rowp) + i));
        } else if (vec8 && i + 3 < P.n_pix) {
          const uchar4 u = __ldcs(reinterpret_cast<const uchar4*>(reinterpret_cast<const uint8_t*>(rowp) + i));
          v = make_float4(float(u.x) / 255.f, float(u.y) / 255.f, float(u.z) / 255.f, float(u.w) / 255.f);
        } else {
          if (i < P.n_pix) v.x = load_pixel<float>(rowp, P.in_dtype, i);
          if (i + 1 < P.n_pix) v.y = load_pixel<float>(rowp, P.in_dtype, i + 1);
          if (i + 2 < P.n_pix) v.z = load_pixel<float>(rowp, P.in_dtype, i + 2);
          if (i + 3 < P.n_pix) v.w = load_pixel<float>(rowp, P.in_dtype, i + 3);
        }
      }
      *reinterpret_cast<float4*>(buf + i) = v;
    }
  }
  __syncwarp();
  const int64_t im = valid ? img : 0;
  row_step<0, false, false>(valid, P, im, buf, buf, ysm, refl, lane);
  row_step<1, false, false>(valid, P, im, buf, buf, ysm, refl, lane);
  row_step<2, false, true>(valid, P, im, buf, ws_out + (size_t)im * kWsStride, ysm, refl, lane);
}

// ---- one ROW step from / to the global workspace (N = 3, 4) ----
// perm (nullable): slot i of the launch works on image perm[i] (-1 = padding);
// n_slots_dev holds the number of slots.  keys_out (nullable): per image, the
// number of vectors this step kept.
template <int N, bool PERM, bool KEYS>
__global__ void __launch_bounds__(kHwWarps * 32, OC_HW_MINB) hw_row_kernel(const __grid_constant__ EncodeParams P,
                                                                  const float* __restrict__ ws_in,
                                                                  float* __restrict__ ws_out,
                                                                  const int32_t* __restrict__ perm,
                                                                  const int32_t* __restrict__ n_slots_dev,
                                                                  uint8_t* __restrict__ keys_out) {
  extern __shared__ __align__(16) float smem_f[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int h = lane >> 4;
  const int64_t slot = ((int64_t)blockIdx.x * kHwWarps + warp) * 2 + h;
  int64_t img = slot;
  if constexpr (PERM) img = slot < *n_slots_dev ? perm[slot] : -1;
  const bool valid = img >= 0 && img < P.N;
  const int64_t im = valid ? img : 0;
  row_step<N, true, true, KEYS>(valid, P, im, ws_in + (size_t)im * kWsStride, ws_out + (size_t)im * kWsStride,
                                smem_f + (size_t)(warp * 2 + h) * ysm_floats(N),
                                smem_f + (size_t)kHwWarps * 2 * ysm_floats(N) +
                                    (size_t)(warp * 2 + h) * refl_floats(4 - N),
                                lane, keys_out);
}

// ---- pairing sort: images with the same number of live rows share a warp ----
// bins[0..31] counts, bins[32..63] starts (even-padded), bins[64..95] cursors, bins[96] = total slots
__global__ void hw_sort_hist_kernel(const uint8_t* __restrict__ keys, int64_t n, int32_t* __restrict__ bins) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicAdd(&bins[keys[i] & (kSortBins - 1)], 1);
}
__global__ void hw_sort_prefix_kernel(int32_t* __restrict__ bins, int32_t* __restrict__ perm) {
  if (threadIdx.x == 0) {
    int start = 0;
    // descending key: the longest lines are scheduled first
    for (int b = kSortBins - 1; b >= 0; --b) {
      const int c = bins[b];
      bins[kSortBins + b] = start;
      bins[2 * kSortBins + b] = 0;
      start += c;
      if (c & 1) perm[start++] = -1;  // pad the bin to an even number of slots
    }
    bins[3 * kSortBins] = start;
  }
}
__global__ void hw_sort_scatter_kernel(const uint8_t* __restrict__ keys, int64_t n, int32_t* __restrict__ bins,
                                       int32_t* __restrict__ perm) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const int b = keys[i] & (kSortBins - 1);
    const int pos = atomicAdd(&bins[2 * kSortBins + b], 1);
    perm[bins[kSortBins + b] + pos] = (int32_t)i;
  }
}

// ---- step 5 (COL): M_5[(s,l), c] = Psi_5[l][s*16 + c], 16 columns of length 64 ----
__global__ void __launch_bounds__(kHwWarps * 32, OC_HW_MINB) hw_col5_kernel(const __grid_constant__ EncodeParams P,
                                                                   const float* __restrict__ ws_in,
                                                                   float* __restrict__ ws_out) {
  constexpr int LOGLP = 1;
  extern __shared__ __align__(16) float smem_f[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int l16 = lane & 15, h = lane >> 4;
  float* refl = smem_f + (size_t)(warp * 2 + h) * refl_floats(LOGLP);
  const int64_t img = ((int64_t)blockIdx.x * kHwWarps + warp) * 2 + h;
  const bool valid = img < P.N;
  const int64_t im = valid ? img : 0;
  const Out io = make_out(P, im, 5);
  const float* psi = ws_in + (size_t)im * kWsStride;  // [32][32]
  const int slot = l16 >> 1, s = l16 & 1;
  u64 T[16], S[16];
  {
    const float* pc = psi + s * 16 + 2 * slot;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float2 v0 = make_float2(0.f, 0.f), v1 = v0;
      if (valid) {
        v0 = __ldg(reinterpret_cast<const float2*>(pc + (2 * j) * 32));
        v1 = __ldg(reinterpret_cast<const float2*>(pc + (2 * j + 1) * 32));
      }
      T[j] = pack2(v0.x, v1.x);
      S[j] = pack2(v0.y, v1.y);
    }
  }
  float nT = dot16<LOGLP>(T, T), nS = dot16<LOGLP>(S, S);
  const float fro = half_sum(s == 0 ? nT + nS : 0.f);
  const float thr2 = kFloor2 * fro;
  const unsigned fmask = __ballot_sync(FULL, fro > 0.f);
  const int mlive = fro > 0.f ? 8 : 0;
  const int mlive_max = fmask ? 8 : 0;
  int sweeps = 0;
  if (mlive_max > 0) {
#if OC_HW_LIFT
    sweeps = jacobi16_lift<LOGLP>(T, S, nT, nS, thr2, lane, mlive, mlive_max, refl);
#else
    // rows l >= rank(bond 5) of Psi_5 are zero (<= 24 for 784 pixels): when the
    // last 4 packed pairs (l >= 24) are zero in the whole warp, skip them
    bool tailnz = false;
#pragma unroll
    for (int j = 12; j < 16; ++j) tailnz = tailnz || (T[j] | S[j]) != 0ull;
    if (__any_sync(FULL, tailnz))
      sweeps = jacobi16<LOGLP>(T, S, nT, nS, thr2, lane, mlive, mlive_max);
    else
      sweeps = jacobi16<LOGLP, 12>(T, S, nT, nS, thr2, lane, mlive, mlive_max);
#endif
    nT = dot16<LOGLP>(T, T);
    nS = dot16<LOGLP>(S, S);
  }
  int rT, rS;
  rank16<LOGLP>(nT, nS, lane, rT, rS);
  const int r_rt = io.r_rt, b_rt = io.b_rt;
  const bool kT = nT > thr2 && rT < r_rt;
  const bool kS = nS > thr2 && rS < r_rt;
  if (valid) {
    if (io.sv) {
      if (s == 0) {
        if (rT < io.sv_width) io.sv[rT] = nT > thr2 ? sqrtf(nT) : 0.f;
        if (rS < io.sv_width) io.sv[rS] = nS > thr2 ? sqrtf(nS) : 0.f;
      }
      for (int k = 16 + l16; k < io.sv_width; k += 16) io.sv[k] = 0.f;
    }
    if (io.sweeps && l16 == 0) *io.sweeps = sweeps;
  }
  // U = normalised columns
  const float iT = kT ? rsqrtf(nT) : 0.f;
  const float iS = kS ? rsqrtf(nS) : 0.f;
  const u64 iT2 = pack2(iT, iT), iS2 = pack2(iS, iS);
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    T[j] = mul2(T[j], iT2);
    S[j] = mul2(S[j], iS2);
  }
  // A_5[(s*b + l), rank]
  if (valid) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float t0, t1, s0, s1;
      unpack2(T[j], t0, t1);
      unpack2(S[j], s0, s1);
      const int l0 = 2 * j, l1 = 2 * j + 1;
      if (l0 < b_rt) {
        if (rT < r_rt) io.A[(s * b_rt + l0) * r_rt + rT] = t0;
        if (rS < r_rt) io.A[(s * b_rt + l0) * r_rt + rS] = s0;
      }
      if (l1 < b_rt) {
        if (rT < r_rt) io.A[(s * b_rt + l1) * r_rt + rT] = t1;
        if (rS < r_rt) io.A[(s * b_rt + l1) * r_rt + rS] = s1;
      }
    }
  }
  // Psi_6[r][c'] = sum_{s,l} U[(s,l)][r] Psi_5[l][s*16 + c']
  u64 aT[8], aS[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) aT[q] = aS[q] = 0ull;
  {
    const float* pm = psi + s * 16;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float t0, t1, s0, s1;
      unpack2(T[j], t0, t1);
      unpack2(S[j], s0, s1);
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        const u64 ut = p ? pack2(t1, t1) : pack2(t0, t0);
        const u64 us = p ? pack2(s1, s1) : pack2(s0, s0);
        const float* row = pm + (2 * j + p) * 32;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
          if (valid) m = ldg4(row + 4 * q4);
          const u64 m0 = pack2(m.x, m.y), m1 = pack2(m.z, m.w);
          aT[2 * q4] = fma2(ut, m0, aT[2 * q4]);
          aT[2 * q4 + 1] = fma2(ut, m1, aT[2 * q4 + 1]);
          aS[2 * q4] = fma2(us, m0, aS[2 * q4]);
          aS[2 * q4 + 1] = fma2(us, m1, aS[2 * q4 + 1]);
        }
      }
    }
  }
  // sum the two s-halves (lane pairs), then lane s = 0 writes row rT, lane s = 1 row rS
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    aT[q] = add2(aT[q], shfl2(aT[q], lane ^ 1));
    aS[q] = add2(aS[q], shfl2(aS[q], lane ^ 1));
  }
  if (valid) {
    float* out = ws_out + (size_t)im * kWsStride;
    const int r = s ? rS : rT;
    if (r < 16) {
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        float4 v;
        const u64 a0 = s ? aS[2 * q4] : aT[2 * q4], a1 = s ? aS[2 * q4 + 1] : aT[2 * q4 + 1];
        unpack2(a0, v.x, v.y);
        unpack2(a1, v.z, v.w);
        *reinterpret_cast<float4*>(out + r * 16 + 4 * q4) = v;
      }
    }
  }
}

inline bool supported(const EncodeParams& P) { return P.L == 10 && P.D > 16; }

// steps 0..5 with the half-warp kernels, 6..9 with the v3 tail
inline int& sort_pairs() {
  static int v = 1;
  return v;
}
inline cudaError_t launch_chain(const EncodeParams& P, float* ws0, float* ws1, void* sort_ws, int64_t chunk,
                                int sms, cudaStream_t st) {
  const int per_cta = kHwWarps * 2;
  const int grid = (int)((P.N + per_cta - 1) / per_cta);
  const size_t smem_head = (size_t)per_cta * (1024 + ysm_floats(2) + refl_floats(3)) * sizeof(float);
  const size_t smem3 = (size_t)per_cta * (ysm_floats(3) + refl_floats(1)) * sizeof(float);
  const size_t smem4 = (size_t)per_cta * (ysm_floats(4) + refl_floats(0)) * sizeof(float);
  const size_t smem5 = (size_t)per_cta * refl_floats(1) * sizeof(float);
  uint8_t* keys = reinterpret_cast<uint8_t*>(sort_ws);
  int32_t* perm = reinterpret_cast<int32_t*>(keys + ((chunk + 255) & ~(int64_t)255));
  int32_t* bins = perm + chunk + 2 * kSortBins;
  const bool sorted = sort_pairs() != 0;
  EncodeEvents& ee = encode_events();
  ee.mark(0, st);
  hw_head_kernel<<<grid, kHwWarps * 32, smem_head, st>>>(P, ws0);
  ee.mark(1, st);
  if (sorted)
    hw_row_kernel<3, false, true><<<grid, kHwWarps * 32, smem3, st>>>(P, ws0, ws1, nullptr, nullptr, keys);
  else
    hw_row_kernel<3, false, false><<<grid, kHwWarps * 32, smem3, st>>>(P, ws0, ws1, nullptr, nullptr, nullptr);
  ee.mark(2, st);
  if (sorted) {
    // step 4 costs (live rows) x (sweeps) per WARP: pair images with equal numbers of live rows
    cudaMemsetAsync(bins, 0, (size_t)4 * kSortBins * sizeof(int32_t), st);
    const int tb = 256, gb = (int)((P.N + tb - 1) / tb);
    hw_sort_hist_kernel<<<gb, tb, 0, st>>>(keys, P.N, bins);
    hw_sort_prefix_kernel<<<1, 32, 0, st>>>(bins, perm);
    hw_sort_scatter_kernel<<<gb, tb, 0, st>>>(keys, P.N, bins, perm);
    const int grid4 = (int)((P.N + kSortBins + per_cta - 1) / per_cta);
    hw_row_kernel<4, true, false><<<grid4, kHwWarps * 32, smem4, st>>>(P, ws1, ws0, perm, bins + 3 * kSortBins, nullptr);
  } else {
    hw_row_kernel<4, false, false><<<grid, kHwWarps * 32, smem4, st>>>(P, ws1, ws0, nullptr, nullptr, nullptr);
  }
  ee.mark(3, st);
  hw_col5_kernel<<<grid, kHwWarps * 32, smem5, st>>>(P, ws0, ws1);
  ee.mark(4, st);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  e = launch_steps<32, 6, 10>(P, ws1, nullptr, sms, st);
  ee.mark(5, st);
  return e;
}

}  // namespace hw
}  // namespace oc
