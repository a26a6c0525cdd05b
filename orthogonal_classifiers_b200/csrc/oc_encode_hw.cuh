// Register-resident encoder kernels (fp32, L = 10, 8 <= D <= 32), steps 4..9 of the
// chain and the older kernels for steps 0..3 (the shipped front end is
// oc_encode_gram.cuh).  Step-major: one kernel per step over all images, Psi
// passing through a 4 KB / image global workspace.
//
//   shipped          step  what
//   hw_row4q_kernel   4    M_4 (2k non-zero rows of 32): Gram-Schmidt on the rows in
//                          registers -> small factor L (2k x 2k); one-sided Jacobi on
//                          the columns of L; U = normalised columns, Psi_5 = U^T M_4.
//                          k lanes per image, floor(32 / k) images of equal k per warp
//                          (device-side counting sort).
//   hw_col5q_kernel   5    M_5 (64 x 16): column-pivoted Gram-Schmidt -> R (16 x 16);
//                          Jacobi on the rows of R, converged vectors = Psi_6;
//                          U = M_5 V / sigma + Gram-Schmidt polish.  16 lanes per image.
//   hw_col_kernel<6>, <7>  Jacobi on the columns, 4 / 2 lanes per image
//   hw_tail89_kernel  8-9  one thread per image
//
//   behind options (oc_set_option; GPU tests compare them with the shipped ones):
//   hw_head_kernel, hw_row_kernel<3>   steps 0-2 / 3 by Jacobi on the image unfoldings
//   hw_row_kernel<4>, hw_row4p_kernel  step 4 by Jacobi on the rows of M_4 (16 lanes per
//                          image / packed), U = M W^T S^-2 + polish
//   hw_col_kernel<5>       step 5 by Jacobi on the 64-element columns
//
// Common machinery: every lane holds the two vectors (T, S) of its slot in
// registers as packed pairs (fma.rn.f32x2); odd-even ordering (round A rotates
// (T_k, S_k), round B (S_k, T_{k+1}) after moving T one slot down and back; every
// rotation swaps positions); cached squared norms, refreshed each sweep;
// branch-free rotation parameters; a `hold` coefficient set makes a lane's
// rotation an exact no-op, so results never depend on the images sharing a warp.
#pragma once
#include "oc_encode_fast.cuh"

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
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <bool G>
__device__ __forceinline__ float4 ld4(const float* p) {
  if constexpr (G) return ldg4(p);
  else return lds4(p);
}

// dot product of two 32-element register vectors, summed over the LP lanes of a slot
// (NE < 16: only the first NE packed pairs can be non-zero)
template <int LOGLP, int NE = 16, int NA, int NB>
__device__ __forceinline__ float dot16(const u64 (&a)[NA], const u64 (&b)[NB]) {
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
// KEEPA picks which products are formed first, i.e. whose registers the results can reuse: false - a' is
// built on c*y in fresh registers and b' overwrites y (right when a moves on by shuffle and b stays: round
// A); true - a' overwrites x and b' is built on c*x in fresh registers (round B: a stays, b is shuffled
// back).  With one form for both, ptxas ended every round pair with ~2 NE register moves (17 % of the
// loop's instructions at NE = 13).
template <int LOGLP, int NE = 16, bool KEEPA = false, int NA>
__device__ __forceinline__ void rotate16(u64 (&a)[NA], u64 (&b)[NA], float& na, float& nb, float thr2,
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
  // A norm update that cancelled leaves a stale cached norm until the refresh at the start of the
  // next sweep.  It can only happen in a rotation by a large angle, which forces that next sweep
  // (`again`), and a stale norm makes a later rotation of the same sweep less effective, never
  // wrong (c and s stay consistent).  Refreshing inside the sweep (OC_TRACK_LOST) costs 2-3 % of
  // the encoder and changes neither the sweep counts nor the results (profiles/encode_r1.md).
#ifdef OC_TRACK_LOST
  lostf = lostf || (rot && (n_second < kCancel * aa || n_first < kCancel * bb));
#endif
  const u64 c2 = pack2(c, c), s2 = pack2(s, s), ns2 = pack2(ns, ns);
#pragma unroll
  for (int e = 0; e < NE; ++e) {
    const u64 x = a[e], y = b[e];
    if constexpr (KEEPA) {
      b[e] = fma2(ns2, y, mul2(c2, x));
      a[e] = fma2(c2, y, mul2(s2, x));
    } else {
      a[e] = fma2(s2, x, mul2(c2, y));
      b[e] = fma2(c2, x, mul2(ns2, y));
    }
  }
  na = n_first;
  nb = n_second;
}

// One-sided Jacobi, odd-even ordering, on the 2*M vectors of each GROUP of
// G = 2^LOGG lanes (one image per group; G = 16 for steps 0..5).
// mlive: live slots of THIS group (0 = nothing to do); mlive_max: warp maximum.
template <int LOGLP, int NE = 16, int LOGG = 4>
__device__ __forceinline__ int jacobi16(u64 (&T)[16], u64 (&S)[16], float& nT, float& nS, float thr2,
                                        int lane, int mlive, int mlive_max) {
  constexpr int LP = 1 << LOGLP;
  constexpr int G = 1 << LOGG;
  constexpr int M = G >> LOGLP;
  const int lg = lane & (G - 1), gb = lane & ~(G - 1);
  const int slot = lg >> LOGLP;
  const int src_dn = gb | ((lg + LP) & (G - 1));
  const int src_up = gb | ((lg - LP) & (G - 1));
  const bool idleB = slot >= mlive - 1;
  const unsigned gmask = (G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << gb;
  int sweeps = 0;
  bool done = mlive == 0;
#pragma unroll 1
  for (int sw = 0; sw < kMaxSweeps; ++sw) {
    bool again = false;
    bool lostf = false;
    {
      const float fT = dot16<LOGLP, NE>(T, T), fS = dot16<LOGLP, NE>(S, S);
      if (!done) {
        nT = fT;
        nS = fS;
      }
    }
#pragma unroll 1
    for (int rr = 0; rr < mlive_max; ++rr) {
      const bool off = done || rr >= mlive;  // this group has nothing left to do in this round
      rotate16<LOGLP, NE>(T, S, nT, nS, thr2, off, lostf, again);
      if (M > 1) {
#pragma unroll
        for (int e = 0; e < NE; ++e) T[e] = shfl2(T[e], src_dn);
        float nR = __shfl_sync(FULL, nT, src_dn);
        rotate16<LOGLP, NE, true>(S, T, nS, nR, thr2, off || idleB, lostf, again);
#pragma unroll
        for (int e = 0; e < NE; ++e) T[e] = shfl2(T[e], src_up);
        nT = __shfl_sync(FULL, nR, src_up);
      }
      {
#ifdef OC_TRACK_LOST
        const unsigned lostmask = __ballot_sync(FULL, lostf);
#else
        const unsigned lostmask = 0u;
#endif
        if (lostmask) {  // warp-uniform, rare after the first sweeps
          const float fT = dot16<LOGLP, NE>(T, T), fS = dot16<LOGLP, NE>(S, S);
          if (lostmask & gmask) {  // only the group that cancelled takes the refreshed norms
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

// One odd-even sweep of Gram-Schmidt projections on the columns of U held by
// the sub == 0 lane of every slot (B2 packed pairs each, unit norm up to the
// error being removed).  Of every pair that meets, the column of the SMALLER
// singular value (key) is orthogonalised against the other; nothing else
// changes, so the well-determined columns keep their accuracy.  Keys and rank
// tags travel with the columns (positions swap as in the Jacobi ordering).
template <int LOGLP, int B2, bool SPREAD = false, bool KEEPA = false>
__device__ __forceinline__ void polish_pair(u64 (&a)[B2], u64 (&b)[B2], float& ka, float& kb, int& ra, int& rb,
                                            bool hold) {
  u64 acc0 = mul2(a[0], b[0]), acc1 = 0ull;
  if (B2 > 1) acc1 = mul2(a[1], b[1]);
#pragma unroll
  for (int e = 2; e < B2; ++e) {
    if (e & 1) acc1 = fma2(a[e], b[e], acc1);
    else acc0 = fma2(a[e], b[e], acc0);
  }
  float g = hsum2(add2(acc0, acc1));
  if constexpr (SPREAD) {  // the columns are spread over the LP lanes of the slot
#pragma unroll
    for (int o = 0; o < LOGLP; ++o) g += __shfl_xor_sync(FULL, g, 1 << o);
  }
  // a' = s x + c y,  b' = c x + ns y  (swap);  hold: identity
  const bool a_big = ka >= kb;
  const float c = hold ? 0.f : 1.f;
  const float s = hold ? 1.f : (a_big ? -g : 0.f);
  const float ns = hold ? 1.f : (a_big ? 0.f : -g);
  const u64 c2 = pack2(c, c), s2 = pack2(s, s), ns2 = pack2(ns, ns);
#pragma unroll
  for (int e = 0; e < B2; ++e) {
    const u64 x = a[e], y = b[e];
    if constexpr (KEEPA) {  // see rotate16
      b[e] = fma2(ns2, y, mul2(c2, x));
      a[e] = fma2(c2, y, mul2(s2, x));
    } else {
      a[e] = fma2(s2, x, mul2(c2, y));
      b[e] = fma2(c2, x, mul2(ns2, y));
    }
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

template <int LOGLP, int B2, bool SPREAD = false>
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
    polish_pair<LOGLP, B2, SPREAD>(T, S, kT, kS, rT, rS, off);
    if (M > 1) {
#pragma unroll
      for (int e = 0; e < B2; ++e) T[e] = shfl2(T[e], src_dn);
      float kR = __shfl_sync(FULL, kT, src_dn);
      int rR = __shfl_sync(FULL, rT, src_dn);
      polish_pair<LOGLP, B2, SPREAD, true>(S, T, kS, kR, rS, rR, off || idleB);
#pragma unroll
      for (int e = 0; e < B2; ++e) T[e] = shfl2(T[e], src_up);
      kT = __shfl_sync(FULL, kR, src_up);
      rT = __shfl_sync(FULL, rR, src_up);
    }
  }
}

// stable descending rank of this lane's two vectors among the NV = 2*M of its half
template <int LOGLP, int LOGG = 4>
__device__ __forceinline__ void rank16(float nT, float nS, int lane, int& rT, int& rS) {
  constexpr int G = 1 << LOGG;
  constexpr int NV = 2 * (G >> LOGLP);
  const int hb = lane & ~(G - 1), slot = (lane & (G - 1)) >> LOGLP;
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

template <int LOGG = 4>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = (1 << LOGG) >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
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
                                         float* psi_out, float* ysm, int lane,
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
  const float fro = group_sum(sub == 0 ? nT + nS : 0.f);
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
    sweeps = jacobi16<LOGLP>(T, S, nT, nS, thr2, lane, mlive, mlive_max);
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
  row_step<0, false, false>(valid, P, im, buf, buf, ysm, lane);
  row_step<1, false, false>(valid, P, im, buf, buf, ysm, lane);
  row_step<2, false, true>(valid, P, im, buf, ws_out + (size_t)im * kWsStride, ysm, lane);
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
                                smem_f + (size_t)(warp * 2 + h) * ysm_floats(N), lane, keys_out);
}

// ---- pairing sort: images with the same number of live rows share a warp ----
// bins[0..31] counts, bins[32..63] starts (even-padded), bins[64..95] cursors, bins[96] = total slots
__global__ void hw_sort_hist_kernel(const uint8_t* __restrict__ keys, int64_t n, int32_t* __restrict__ bins) {
  // block-local histogram in shared memory, one global atomic per non-empty bin and block
  // (the keys take ~7 values: 70,000 global atomics on 7 addresses serialise)
  __shared__ int sh[kSortBins];
  if (threadIdx.x < kSortBins) sh[threadIdx.x] = 0;
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicAdd(&sh[keys[i] & (kSortBins - 1)], 1);
  __syncthreads();
  if (threadIdx.x < kSortBins && sh[threadIdx.x]) atomicAdd(&bins[threadIdx.x], sh[threadIdx.x]);
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

// ---- packed step 4 ----------------------------------------------------------
// Step 4 rotates the 2 * k non-zero rows of M_4 (k = rows the previous step kept,
// 7..13 for MNIST-like images), i.e. k slots: with 16 lanes per image 3..9 lanes
// idle.  The packed kernel gives an image exactly k lanes and puts floor(32 / k)
// (<= 4) images of EQUAL k into a warp (the counting sort below makes the lists):
// 2.5 images per warp on average instead of 2, same instruction count per warp.
// Results do not depend on the warp's other images (hold = exact no-op).

// One-sided Jacobi, odd-even ordering, on the 2 * gsz vectors of each lane group
// [gbase, gbase + gsz), one lane per slot, NE packed pairs per vector; groups of
// different size side by side (gsz need not be a power of two: the line does not
// wrap, the lanes at its ends idle in round B).
template <int NE, int NA>
__device__ __forceinline__ int jacobi_ring(u64 (&T)[NA], u64 (&S)[NA], float& nT, float& nS, float thr2,
                                           int lane, int gsz, int gbase, int mlive, int mlive_max) {
  const int slot = lane - gbase;
  const int src_dn = slot + 1 < gsz ? lane + 1 : gbase;
  const int src_up = slot > 0 ? lane - 1 : gbase + gsz - 1;
  const bool idleB = slot >= mlive - 1;
  const unsigned gmask = (gsz >= 32 ? 0xffffffffu : ((1u << gsz) - 1u)) << gbase;
  int sweeps = 0;
  bool done = mlive == 0;
#pragma unroll 1
  for (int sw = 0; sw < kMaxSweeps; ++sw) {
    bool again = false;
    bool lostf = false;
    {
      const float fT = dot16<0, NE>(T, T), fS = dot16<0, NE>(S, S);
      if (!done) {
        nT = fT;
        nS = fS;
      }
    }
#pragma unroll 1  // (unrolled twice the loop loses ~4 % of its instructions and 7 % of its speed: instruction cache)
    for (int rr = 0; rr < mlive_max; ++rr) {
      const bool off = done || rr >= mlive;
#ifndef OC_RING_QP
#define OC_RING_QP(ne) ((ne) == 16)
#endif
      // which round forms its products on which operand (rotate16<KEEPA>) decides how many register moves
      // ptxas leaves at the end of the round pair; the better assignment differs between instantiations
      // (26-element loop: 21 moves with (false, true), 32-element loop: 71 with that and 19 with (true, false))
      rotate16<0, NE, OC_RING_QP(NE)>(T, S, nT, nS, thr2, off, lostf, again);
#pragma unroll
      for (int e = 0; e < NE; ++e) T[e] = shfl2(T[e], src_dn);
      float nR = __shfl_sync(FULL, nT, src_dn);
      rotate16<0, NE, !OC_RING_QP(NE)>(S, T, nS, nR, thr2, off || idleB, lostf, again);
#pragma unroll
      for (int e = 0; e < NE; ++e) T[e] = shfl2(T[e], src_up);
      nT = __shfl_sync(FULL, nR, src_up);
      {
#ifdef OC_TRACK_LOST
        const unsigned lostmask = __ballot_sync(FULL, lostf);
#else
        const unsigned lostmask = 0u;
#endif
        if (lostmask) {
          const float fT = dot16<0, NE>(T, T), fS = dot16<0, NE>(S, S);
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

// polish16 for lane groups [gbase, gbase + gsz), one lane per slot
template <int B2>
__device__ __forceinline__ void polish_ring(u64 (&T)[B2], u64 (&S)[B2], float& kT, float& kS, int& rT, int& rS,
                                            int lane, int gsz, int gbase, int mlive, int mlive_max) {
  const int slot = lane - gbase;
  const int src_dn = slot + 1 < gsz ? lane + 1 : gbase;
  const int src_up = slot > 0 ? lane - 1 : gbase + gsz - 1;
  const bool idleB = slot >= mlive - 1;
#pragma unroll 1
  for (int rr = 0; rr < mlive_max; ++rr) {
    const bool off = rr >= mlive;
    polish_pair<0, B2>(T, S, kT, kS, rT, rS, off);
#pragma unroll
    for (int e = 0; e < B2; ++e) T[e] = shfl2(T[e], src_dn);
    float kR = __shfl_sync(FULL, kT, src_dn);
    int rR = __shfl_sync(FULL, rT, src_dn);
    polish_pair<0, B2, false, true>(S, T, kS, kR, rS, rR, off || idleB);
#pragma unroll
    for (int e = 0; e < B2; ++e) T[e] = shfl2(T[e], src_up);
    kT = __shfl_sync(FULL, kR, src_up);
    rT = __shfl_sync(FULL, rR, src_up);
  }
}

// perm4: four image indices per warp (-1 = empty); bins[3 * kSortBins] = number of warps.
__global__ void __launch_bounds__(kHwWarps * 32, OC_HW_MINB)
hw_row4p_kernel(const __grid_constant__ EncodeParams P, const float* __restrict__ ws_in, float* __restrict__ ws_out,
                const int32_t* __restrict__ perm4, const int32_t* __restrict__ bins,
                const uint8_t* __restrict__ keys) {
  constexpr int B = 16, Q = 32, YSTR = 2 * B + 2;
  extern __shared__ __align__(16) float smem_f[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t w = (int64_t)blockIdx.x * kHwWarps + warp;
  if (w >= bins[3 * kSortBins]) return;  // per-warp shared memory only: no block barrier below
  const int4 ent = __ldg(reinterpret_cast<const int4*>(perm4) + w);
  const int k0 = ent.x >= 0 ? (int)keys[ent.x] : 0;
  const int gs = k0 > 1 ? k0 : 1;                 // lanes (slots) per image
  const int ipw = 32 / gs < 4 ? 32 / gs : 4;      // images of this warp
  const int idx = lane / gs;
  int img = idx == 0 ? ent.x : (idx == 1 ? ent.y : (idx == 2 ? ent.z : ent.w));
  if (idx >= ipw) img = -1;
  const bool valid = img >= 0 && img < P.N;
  const int gsz = idx < ipw ? gs : 1;
  const int gbase = idx < ipw ? idx * gs : lane;
  const int slot = lane - gbase;
  const int64_t im = valid ? img : 0;
  const float* psi_in = ws_in + (size_t)im * kWsStride;
  float* psi_out = ws_out + (size_t)im * kWsStride;
  float* ysm = smem_f + (size_t)warp * 64 * YSTR;

  u64 T[16], S[16];
  {
    const float* pt = psi_in + slot * 2 * Q;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f), s = t;
      if (valid) {
        t = ldg4(pt + 4 * i);
        s = ldg4(pt + Q + 4 * i);
      }
      T[2 * i] = pack2(t.x, t.y); T[2 * i + 1] = pack2(t.z, t.w);
      S[2 * i] = pack2(s.x, s.y); S[2 * i + 1] = pack2(s.z, s.w);
    }
  }
  float nT = dot16<0>(T, T), nS = dot16<0>(S, S);
  float fro = 0.f;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float v = __shfl_sync(FULL, nT + nS, gbase + j);
    if (j < gsz) fro += v;
  }
  const float thr2 = kFloor2 * fro;
  const unsigned livemask = __ballot_sync(FULL, (nT + nS) > 0.f);
  const unsigned mine = (livemask >> gbase) & ((1u << gsz) - 1u);
  const int mlive = mine ? 32 - __clz(mine) : 0;
  const int mlive_max = __reduce_max_sync(FULL, mlive);
  int sweeps = 0;
  if (mlive_max > 0) {
    sweeps = jacobi_ring<16>(T, S, nT, nS, thr2, lane, gsz, gbase, mlive, mlive_max);
    nT = dot16<0>(T, T);
    nS = dot16<0>(S, S);
  }
  // stable descending rank among the 2 * gsz vectors of the image
  int rT = 0, rS = 0;
  {
    const int pT = 2 * slot, pS = pT + 1;
#pragma unroll 8
    for (int j = 0; j < 32; ++j) {
      const float v = __shfl_sync(FULL, (j & 1) ? nS : nT, gbase + (j >> 1));
      if (j < 2 * gsz) {
        rT += (v > nT) || (v == nT && j < pT);
        rS += (v > nS) || (v == nS && j < pS);
      }
    }
  }
  const Out io = make_out(P, im, 4);
  const int r_rt = io.r_rt, b_rt = io.b_rt;
  const bool kT = nT > thr2 && rT < r_rt;
  const bool kS = nS > thr2 && rS < r_rt;
  if (valid) {
    if (io.sv) {
      if (rT < io.sv_width) io.sv[rT] = nT > thr2 ? sqrtf(nT) : 0.f;
      if (rS < io.sv_width) io.sv[rS] = nS > thr2 ? sqrtf(nS) : 0.f;
      for (int k = 2 * gsz + slot; k < io.sv_width; k += gsz) io.sv[k] = 0.f;
    }
    if (io.sweeps && slot == 0) *io.sweeps = sweeps;
  }
  // ---- U[(l,s), k] = (1/n_k) sum_c Psi_4[l][s*Q + c] W_k[c]  -> shared, column-major ----
  float* yT = ysm + lane * YSTR;
  float* yS = ysm + (32 + lane) * YSTR;
  {
    const float iT = kT ? 1.f / nT : 0.f;
    const float iS = kS ? 1.f / nS : 0.f;
#pragma unroll 1
    for (int l = 0; l < B; ++l) {
      float aT[2] = {0.f, 0.f}, aS[2] = {0.f, 0.f};
      if (l < mlive_max) {  // warp-uniform; rows l >= mlive are zero
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const float* mr = psi_in + l * 2 * Q + s * Q;
          u64 t0 = 0ull, t1 = 0ull, s0 = 0ull, s1 = 0ull;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) m = ldg4(mr + 4 * i);
            const u64 m0 = pack2(m.x, m.y), m1 = pack2(m.z, m.w);
            t0 = fma2(m0, T[2 * i], t0);
            t1 = fma2(m1, T[2 * i + 1], t1);
            s0 = fma2(m0, S[2 * i], s0);
            s1 = fma2(m1, S[2 * i + 1], s1);
          }
          aT[s] = hsum2(add2(t0, t1)) * iT;
          aS[s] = hsum2(add2(s0, s1)) * iS;
        }
      }
      *reinterpret_cast<float2*>(yT + 2 * l) = make_float2(aT[0], aT[1]);
      *reinterpret_cast<float2*>(yS + 2 * l) = make_float2(aS[0], aS[1]);
    }
  }
  // ---- Psi_5: row `rank` <- vector (zero when dropped); rows >= 2 gsz are zero ----
  if (valid) {
    float* pT = psi_out + rT * Q;
    float* pS = psi_out + rS * Q;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 t, s;
      unpack2(T[2 * i], t.x, t.y); unpack2(T[2 * i + 1], t.z, t.w);
      unpack2(S[2 * i], s.x, s.y); unpack2(S[2 * i + 1], s.z, s.w);
      if (!kT) t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!kS) s = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(pT + 4 * i) = t;
      *reinterpret_cast<float4*>(pS + 4 * i) = s;
    }
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e = 2 * gsz * (Q / 4) + slot; e < 2 * B * (Q / 4); e += gsz)
      *reinterpret_cast<float4*>(psi_out + 4 * e) = z;
  }
  __syncwarp();
  // ---- polish U (see polish_pair), normalise, A_4[(s*b + l), rank] ----
  {
    u64 UT[B], US[B];
#pragma unroll
    for (int e = 0; e < B; ++e) {
      UT[e] = *reinterpret_cast<const u64*>(yT + 2 * e);
      US[e] = *reinterpret_cast<const u64*>(yS + 2 * e);
    }
    float keyT = kT ? nT : 0.f, keyS = kS ? nS : 0.f;
    int tagT = rT, tagS = rS;
    polish_ring<B>(UT, US, keyT, keyS, tagT, tagS, lane, gsz, gbase, mlive, mlive_max);
    u64 qT = 0ull, qS = 0ull;
#pragma unroll
    for (int e = 0; e < B; ++e) {
      qT = fma2(UT[e], UT[e], qT);
      qS = fma2(US[e], US[e], qS);
    }
    const float jT = keyT > 0.f ? rsqrtf(hsum2(qT)) : 0.f;
    const float jS = keyS > 0.f ? rsqrtf(hsum2(qS)) : 0.f;
    if (valid) {
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
      // columns >= 2 gsz of A_4 belong to vectors this image does not have: zero
      const int c0 = 2 * gsz, nc = r_rt - c0;
      if (nc > 0)
        for (int e = slot; e < 2 * b_rt * nc; e += gsz) io.A[(e / nc) * r_rt + c0 + e % nc] = 0.f;
    }
  }
}

// ---- packed step 4, QR-preconditioned ----------------------------------------
// M_4 (2k rows of 32) = L Q with Q's rows orthonormal: the left singular vectors
// of M_4 are those of the SMALL factor L (2k x 2k), so the Jacobi iteration runs
// on the columns of L - vectors of 2k <= 26 elements instead of 32, 5.1 sweeps
// instead of 7.1 (rows of R = L^T are closer to orthogonal than the rows of M_4:
// Drmac-Veselic preconditioning, here without pivoting, the rows arrive sorted
// by the previous step's singular values) - and U comes out of the iteration
// directly (normalised columns): no back-substitution, no Gram-Schmidt polish.
// L by one pass of modified Gram-Schmidt on the rows in registers (R is backward
// stable even where Q is not orthonormal; Q is not used).  Psi_5 = U^T M_4.
// Same warp lists as hw_row4p_kernel.
constexpr int kQ4NE = 13;   // packed pairs of an L column: 2k <= 26 elements
constexpr int kQ4RS = 36;   // floats per row of R in shared memory (2k <= 32 entries + pad)
// Everything behind the factorisation, for L columns of NE packed pairs (2 gs <= 2 NE elements)
template <int NE>
__device__ __forceinline__ void q4_tail(const EncodeParams& P, bool valid, bool ingrp, int gs, int gsz, int gbase,
                                        int slot, int lane, int64_t im, const float* __restrict__ psi_in,
                                        float* __restrict__ psi_out, const float* __restrict__ Rw, float thr2) {
  constexpr int B = 16, Q = 32;
  // ---- columns (2 slot, 2 slot + 1) of L = rows of R -> registers ----
  u64 T[NE], S[NE];
  {
    const float* r0 = Rw + (size_t)(2 * gbase + 2 * slot) * kQ4RS;
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      float2 a = make_float2(0.f, 0.f), b = a;
      if (ingrp && e < gsz) {
        a = *reinterpret_cast<const float2*>(r0 + 2 * e);
        b = *reinterpret_cast<const float2*>(r0 + kQ4RS + 2 * e);
      }
      T[e] = pack2(a.x, a.y);
      S[e] = pack2(b.x, b.y);
    }
  }
  float nT = dot16<0, NE>(T, T), nS = dot16<0, NE>(S, S);
  const unsigned livemask = __ballot_sync(FULL, (nT + nS) > 0.f);
  const unsigned mine = (livemask >> gbase) & ((1u << gsz) - 1u);
  const int mlive = mine ? 32 - __clz(mine) : 0;
  const int mlive_max = __reduce_max_sync(FULL, mlive);
  int sweeps = 0;
  if (mlive_max > 0) {
    sweeps = jacobi_ring<NE>(T, S, nT, nS, thr2, lane, gsz, gbase, mlive, mlive_max);
    nT = dot16<0, NE>(T, T);
    nS = dot16<0, NE>(S, S);
  }
  int rT = 0, rS = 0;
  {
    const int pT = 2 * slot, pS = pT + 1;
#pragma unroll 8
    for (int j = 0; j < 32; ++j) {
      const float v = __shfl_sync(FULL, (j & 1) ? nS : nT, gbase + (j >> 1));
      if (j < 2 * gsz) {
        rT += (v > nT) || (v == nT && j < pT);
        rS += (v > nS) || (v == nS && j < pS);
      }
    }
  }
  const Out io = make_out(P, im, 4);
  const int r_rt = io.r_rt, b_rt = io.b_rt;
  const bool kT = nT > thr2 && rT < r_rt;
  const bool kS = nS > thr2 && rS < r_rt;
  if (valid) {
    if (io.sv) {
      if (rT < io.sv_width) io.sv[rT] = nT > thr2 ? sqrtf(nT) : 0.f;
      if (rS < io.sv_width) io.sv[rS] = nS > thr2 ? sqrtf(nS) : 0.f;
      for (int k = 2 * gsz + slot; k < io.sv_width; k += gsz) io.sv[k] = 0.f;
    }
    if (io.sweeps && slot == 0) *io.sweeps = sweeps;
  }
  // ---- U = normalised columns: U[2l + s][vector];  A_4[(s*b + l), rank] ----
  {
    const float iT = kT ? rsqrtf(nT) : 0.f, iS = kS ? rsqrtf(nS) : 0.f;
    const u64 iT2 = pack2(iT, iT), iS2 = pack2(iS, iS);
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      T[e] = mul2(T[e], iT2);
      S[e] = mul2(S[e], iS2);
    }
  }
  if (valid) {
#pragma unroll
    for (int l = 0; l < B; ++l) {
      float t0 = 0.f, t1 = 0.f, s0 = 0.f, s1 = 0.f;
      if (l < NE) {
        unpack2(T[l < NE ? l : 0], t0, t1);
        unpack2(S[l < NE ? l : 0], s0, s1);
      }
      if (l < b_rt) {
        if (rT < r_rt) {
          io.A[l * r_rt + rT] = t0;
          io.A[(b_rt + l) * r_rt + rT] = t1;
        }
        if (rS < r_rt) {
          io.A[l * r_rt + rS] = s0;
          io.A[(b_rt + l) * r_rt + rS] = s1;
        }
      }
    }
    const int c0 = 2 * gsz, nc = r_rt - c0;
    if (nc > 0)
      for (int e = slot; e < 2 * b_rt * nc; e += gsz) io.A[(e / nc) * r_rt + c0 + e % nc] = 0.f;
  }
  // ---- Psi_5[rank][c] = sum_i U[i][vector] M_4[i][c], two passes of 16 columns ----
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
    u64 aT[8], aS[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) aT[q] = aS[q] = 0ull;
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      if (e < gs) {  // warp-uniform; rows i >= 2 gs of M_4 (and of U) are zero
        float t0, t1, s0, s1;
        unpack2(T[e], t0, t1);
        unpack2(S[e], s0, s1);
#pragma unroll
        for (int sdx = 0; sdx < 2; ++sdx) {
          const u64 ut = sdx ? pack2(t1, t1) : pack2(t0, t0);
          const u64 us = sdx ? pack2(s1, s1) : pack2(s0, s0);
          const float* mr = psi_in + e * 2 * Q + sdx * Q + 16 * half;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) m = ldg4(mr + 4 * q4);
            const u64 m0 = pack2(m.x, m.y), m1 = pack2(m.z, m.w);
            aT[2 * q4] = fma2(ut, m0, aT[2 * q4]);
            aT[2 * q4 + 1] = fma2(ut, m1, aT[2 * q4 + 1]);
            aS[2 * q4] = fma2(us, m0, aS[2 * q4]);
            aS[2 * q4 + 1] = fma2(us, m1, aS[2 * q4 + 1]);
          }
        }
      }
    }
    if (valid) {
      float* pT = psi_out + rT * Q + 16 * half;
      float* pS = psi_out + rS * Q + 16 * half;
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        float4 t, s;
        unpack2(aT[2 * q4], t.x, t.y); unpack2(aT[2 * q4 + 1], t.z, t.w);
        unpack2(aS[2 * q4], s.x, s.y); unpack2(aS[2 * q4 + 1], s.z, s.w);
        *reinterpret_cast<float4*>(pT + 4 * q4) = t;
        *reinterpret_cast<float4*>(pS + 4 * q4) = s;
      }
    }
  }
  if (valid) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e = 2 * gsz * (Q / 4) + slot; e < 2 * B * (Q / 4); e += gsz)
      *reinterpret_cast<float4*>(psi_out + 4 * e) = z;
  }
}

__global__ void __launch_bounds__(kHwWarps * 32, OC_HW_MINB)
hw_row4q_kernel(const __grid_constant__ EncodeParams P, const float* __restrict__ ws_in, float* __restrict__ ws_out,
                const int32_t* __restrict__ perm4, const int32_t* __restrict__ bins,
                const uint8_t* __restrict__ keys) {
  constexpr int B = 16, Q = 32;
  extern __shared__ __align__(16) float smem_f[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t w = (int64_t)blockIdx.x * kHwWarps + warp;
  if (w >= bins[3 * kSortBins]) return;  // per-warp shared memory only: no block barrier below
  // lane groups of the warp's (up to four) images: k_j lanes each, in list order; the first image has
  // the longest line (hosts before guests, see hw_pack_prefix_kernel), empty list slots take no lanes
  const int4 ent = __ldg(reinterpret_cast<const int4*>(perm4) + w);
  const int k0 = ent.x >= 0 ? max((int)keys[ent.x], 1) : 0;
  const int k1 = ent.y >= 0 ? max((int)keys[ent.y], 1) : 0;
  const int k2 = ent.z >= 0 ? max((int)keys[ent.z], 1) : 0;
  const int k3 = ent.w >= 0 ? max((int)keys[ent.w], 1) : 0;
  const int gs = max(max(max(k0, k1), max(k2, k3)), 1);  // longest line of the warp
  const int b1 = k0, b2 = b1 + k1, b3 = b2 + k2, b4 = b3 + k3;
  const int idx = (lane >= b1) + (lane >= b2) + (lane >= b3);
  const bool ingrp = lane < b4;
  int img = idx == 0 ? ent.x : (idx == 1 ? ent.y : (idx == 2 ? ent.z : ent.w));
  if (!ingrp) img = -1;
  const bool valid = img >= 0 && img < P.N;
  const int gsz = ingrp ? (idx == 0 ? k0 : (idx == 1 ? k1 : (idx == 2 ? k2 : k3))) : 1;
  const int gbase = ingrp ? (idx == 0 ? 0 : (idx == 1 ? b1 : (idx == 2 ? b2 : b3))) : lane;
  const int slot = lane - gbase;
  const int64_t im = valid ? img : 0;
  const float* psi_in = ws_in + (size_t)im * kWsStride;
  float* psi_out = ws_out + (size_t)im * kWsStride;
  float* Rw = smem_f + (size_t)warp * 64 * kQ4RS;  // R[j][i] of the image at row 2 * gbase + j

  float thr2;
  {
    // ---- rows (2 slot, 2 slot + 1) of M_4 -> registers; MGS, natural order ----
    u64 T[16], S[16];
    {
      const float* pt = psi_in + slot * 2 * Q;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f), s = t;
        if (valid) {
          t = ldg4(pt + 4 * i);
          s = ldg4(pt + Q + 4 * i);
        }
        T[2 * i] = pack2(t.x, t.y); T[2 * i + 1] = pack2(t.z, t.w);
        S[2 * i] = pack2(s.x, s.y); S[2 * i + 1] = pack2(s.z, s.w);
      }
    }
    const float n0 = dot16<0>(T, T) + dot16<0>(S, S);
    float fro = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float v = __shfl_sync(FULL, n0, gbase + j);
      if (j < gsz) fro += v;
    }
    thr2 = kFloor2 * fro;
    const int vT = 2 * slot, vS = vT + 1;
    float* rrow = Rw + (size_t)(2 * gbase) * kQ4RS + vT;
#pragma unroll 1
    for (int jj = 0; jj < gs; ++jj) {
      const int owner = gbase + (jj < gsz ? jj : 0);
#pragma unroll
      for (int par = 0; par < 2; ++par) {
        const int j = 2 * jj + par;
        u64 qv[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) qv[e] = shfl2(par ? S[e] : T[e], owner);
        float rT = dot16<0>(qv, T), rS = dot16<0>(qv, S);
        // |q|^2 is the owner's own dot product (its vector IS q): one shuffle instead of a third dot
        const float qn2 = __shfl_sync(FULL, par ? rS : rT, owner);
        const bool livej = qn2 > thr2;
        const float inv2 = livej ? 1.f / qn2 : 0.f, invn = livej ? rsqrtf(qn2) : 0.f;
        if (vT <= j) rT = 0.f;
        if (vS <= j) rS = 0.f;
        const float cT = -rT * inv2, cS = -rS * inv2;
        const u64 cT2 = pack2(cT, cT), cS2 = pack2(cS, cS);
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          T[e] = fma2(cT2, qv[e], T[e]);
          S[e] = fma2(cS2, qv[e], S[e]);
        }
        if (ingrp && jj < gsz) {
          const float dg = qn2 * invn;  // |q_j|: the diagonal of R
          *reinterpret_cast<float2*>(rrow + (size_t)j * kQ4RS) =
              make_float2(vT == j ? dg : rT * invn, vS == j ? dg : rS * invn);
        }
      }
    }
  }
  __syncwarp();
  // vectors of 2 gs elements: the rotation cost is proportional to the padded length
  // (instantiations for 11 and 12: step 4 1.017 -> 0.977 ms per 70,000 images; one for 9 bought nothing)
  if (gs <= 10) q4_tail<10>(P, valid, ingrp, gs, gsz, gbase, slot, lane, im, psi_in, psi_out, Rw, thr2);
#ifndef OC_Q4_COARSE
  else if (gs == 11) q4_tail<11>(P, valid, ingrp, gs, gsz, gbase, slot, lane, im, psi_in, psi_out, Rw, thr2);
  else if (gs == 12) q4_tail<12>(P, valid, ingrp, gs, gsz, gbase, slot, lane, im, psi_in, psi_out, Rw, thr2);
#endif
  else if (gs <= kQ4NE) q4_tail<kQ4NE>(P, valid, ingrp, gs, gsz, gbase, slot, lane, im, psi_in, psi_out, Rw, thr2);
  else q4_tail<16>(P, valid, ingrp, gs, gsz, gbase, slot, lane, im, psi_in, psi_out, Rw, thr2);  // n_pix > 832
}

// counting sort for the packed kernel: bins[0..31] counts, [32..63] first warp of the bin,
// [64..95] cursors, [96] number of warps, [97] number of guest records, [kPackOwn + b] images of bin b
// that stay in the bin's own warps, [kPackRec + 4 r ..] guest record r = (guest bin, images, first host
// warp, list slot).
// Guests (mix != 0): a warp of floor(32 / k) images of k lanes leaves 32 mod k lanes idle - 10 of 32 at
// k = 11, 8 at k = 12.  A warp costs (its longest line) x (its slowest image), so an image with a shorter
// line rides in those lanes for free: every warp of a bin with a free list slot takes one guest from the
// largest bin that fits (the guest that is most expensive to run on its own).  Hosts come first in the
// list, so the first image of a warp has its longest line.  Nothing depends on which images share a warp.
constexpr int kPackOwn = 4 * kSortBins, kPackRec = 5 * kSortBins, kPackMaxRec = 32;
constexpr int kPackInts = kPackRec + 4 * kPackMaxRec;
__global__ void hw_pack_prefix_kernel(int32_t* __restrict__ bins, int mix) {
  __shared__ int rem[kSortBins];  // (a local array indexed at run time lived in local memory: 11 us on one thread)
  if (threadIdx.x < kSortBins) rem[threadIdx.x] = bins[threadIdx.x];
  __syncthreads();
  if (threadIdx.x == 0) {
    int start = 0, nrec = 0;
    for (int b = kSortBins - 1; b >= 0; --b) {  // longest lines first
      const int c = rem[b];
      const int gs = b > 1 ? b : 1;
      const int ipw = 32 / gs < 4 ? 32 / gs : 4;
      const int W = (c + ipw - 1) / ipw;
      bins[kSortBins + b] = start;
      bins[2 * kSortBins + b] = 0;
      bins[kPackOwn + b] = c;
      // (no guests in the 32-element instantiation, b > kQ4NE: its rotations round differently - see OC_RING_QP -
      // and an image's bits must not depend on its warp)
      if (mix && ipw < 4 && b <= kQ4NE) {
        int filled = 0;
        const int freel = 32 - ipw * gs;
        for (int g = freel < b - 1 ? freel : b - 1; g >= 2 && filled < W && nrec < kPackMaxRec; --g) {
          const int n = rem[g] < W - filled ? rem[g] : W - filled;
          if (n <= 0) continue;
          int32_t* rec = bins + kPackRec + 4 * nrec++;
          rec[0] = g;
          rec[1] = n;
          rec[2] = start + filled;
          rec[3] = ipw;
          rem[g] -= n;
          filled += n;
        }
      }
      start += W;
    }
    bins[3 * kSortBins] = start;
    bins[3 * kSortBins + 1] = nrec;
  }
}
__global__ void hw_pack_scatter_kernel(const uint8_t* __restrict__ keys, int64_t n, int32_t* __restrict__ bins,
                                       int32_t* __restrict__ perm4) {
  // positions inside a bin: block-local ranks from shared-memory atomics + one global atomic per
  // non-empty bin and block for the block's base
  __shared__ int sh[kSortBins], base[kSortBins];
  if (threadIdx.x < kSortBins) sh[threadIdx.x] = 0;
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int b = 0, local = 0;
  if (i < n) {
    b = keys[i] & (kSortBins - 1);
    local = atomicAdd(&sh[b], 1);
  }
  __syncthreads();
  if (threadIdx.x < kSortBins && sh[threadIdx.x])
    base[threadIdx.x] = atomicAdd(&bins[2 * kSortBins + threadIdx.x], sh[threadIdx.x]);
  __syncthreads();
  if (i < n) {
    const int gs = b > 1 ? b : 1;
    const int ipw = 32 / gs < 4 ? 32 / gs : 4;
    const int pos = base[b] + local;
    const int own = bins[kPackOwn + b];
    if (pos < own) {
      perm4[4 * (bins[kSortBins + b] + pos / ipw) + pos % ipw] = (int32_t)i;
    } else {  // guest: the bin's records in order
      int q = pos - own;
      const int nrec = bins[3 * kSortBins + 1];
      for (int r = 0; r < nrec; ++r) {
        const int32_t* rec = bins + kPackRec + 4 * r;
        if (rec[0] != b) continue;
        if (q < rec[1]) {
          perm4[4 * (rec[2] + q) + rec[3]] = (int32_t)i;
          break;
        }
        q -= rec[1];
      }
    }
  }
}

// ---- COL steps 5, 6, 7: M_n[(s,l), c] = Psi_n[l][s*Q + c], Q = 2^(9-n) columns of
//      length 2B = 2^(11-n).  Slot k owns columns (2k, 2k+1); one image per
//      group of G lanes:  n  cols x len  slots  lanes/slot  G   images/warp
//                         5   16 x 64      8        2      16       2
//                         6    8 x 32      4        1       4       8
//                         7    4 x 16      2        1       2      16
//      Converged columns are U.S; A_n = U, Psi_{n+1} = U^T M_n.
template <int N>
struct ColShape {
  static constexpr int B = 1 << (10 - N), Q = 1 << (9 - N);
  static constexpr int LOGLP = N == 5 ? 1 : 0, LP = 1 << LOGLP;
  static constexpr int M = Q / 2, G = M * LP, LOGG = clog2(G);
  static constexpr int EL = 2 * B / LP;  // elements per lane and vector (32, 32, 16)
  static constexpr int NE = EL / 2;      // packed pairs
  static constexpr int IPW = 32 / G;     // images per warp
};

template <int N>
__global__ void __launch_bounds__(kHwWarps * 32, OC_HW_MINB) hw_col_kernel(const __grid_constant__ EncodeParams P,
                                                                           const float* __restrict__ ws_in,
                                                                           float* __restrict__ ws_out) {
  using C = ColShape<N>;
  constexpr int B = C::B, Q = C::Q, LOGLP = C::LOGLP, LP = C::LP, M = C::M, G = C::G, LOGG = C::LOGG;
  constexpr int EL = C::EL, NE = C::NE, IPW = C::IPW;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lg = lane & (G - 1), grp = lane >> LOGG;
  const int64_t img = ((int64_t)blockIdx.x * kHwWarps + warp) * IPW + grp;
  const bool valid = img < P.N;
  const int64_t im = valid ? img : 0;
  const Out io = make_out(P, im, N);
  const float* psi = ws_in + (size_t)im * kWsStride;  // [B][2Q]
  const int slot = lg >> LOGLP, sub = lg & (LP - 1);
  // this lane's elements e = sub*EL + i, i < EL  <->  (s, l) = (e / B, e % B);
  // packed pair j holds e = e0 + 2j, e0 + 2j + 1 (same s: B is even)
  const int e0 = sub * EL;
  u64 T[16], S[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) T[j] = S[j] = 0ull;
  {
    const float* pc = psi + 2 * slot;
#pragma unroll
    for (int j = 0; j < NE; ++j) {
      const int e = e0 + 2 * j, sj = e / B, lj = e % B;
      float2 v0 = make_float2(0.f, 0.f), v1 = v0;
      if (valid) {
        v0 = __ldg(reinterpret_cast<const float2*>(pc + lj * 2 * Q + sj * Q));
        v1 = __ldg(reinterpret_cast<const float2*>(pc + (lj + 1) * 2 * Q + sj * Q));
      }
      T[j] = pack2(v0.x, v1.x);
      S[j] = pack2(v0.y, v1.y);
    }
  }
  float nT = dot16<LOGLP, NE>(T, T), nS = dot16<LOGLP, NE>(S, S);
  const float fro = group_sum<LOGG>(sub == 0 ? nT + nS : 0.f);
  const float thr2 = kFloor2 * fro;
  const int mlive = fro > 0.f ? M : 0;
  const int mlive_max = __any_sync(FULL, fro > 0.f) ? M : 0;
  int sweeps = 0;
  if (mlive_max > 0) {
    if constexpr (N == 5) {
      // rows l >= rank(bond 5) of Psi_5 are zero (<= 24 for 784 pixels): when the
      // last 4 packed pairs (l >= 24) are zero in the whole warp, skip them
      bool tailnz = false;
#pragma unroll
      for (int j = 12; j < 16; ++j) tailnz = tailnz || (T[j] | S[j]) != 0ull;
      if (__any_sync(FULL, tailnz))
        sweeps = jacobi16<LOGLP, 16, LOGG>(T, S, nT, nS, thr2, lane, mlive, mlive_max);
      else
        sweeps = jacobi16<LOGLP, 12, LOGG>(T, S, nT, nS, thr2, lane, mlive, mlive_max);
    } else {
      sweeps = jacobi16<LOGLP, NE, LOGG>(T, S, nT, nS, thr2, lane, mlive, mlive_max);
    }
    nT = dot16<LOGLP, NE>(T, T);
    nS = dot16<LOGLP, NE>(S, S);
  }
  int rT, rS;
  rank16<LOGLP, LOGG>(nT, nS, lane, rT, rS);
  const int r_rt = io.r_rt, b_rt = io.b_rt;
  const bool kT = nT > thr2 && rT < r_rt;
  const bool kS = nS > thr2 && rS < r_rt;
  if (valid) {
    if (io.sv) {
      if (sub == 0) {
        if (rT < io.sv_width) io.sv[rT] = nT > thr2 ? sqrtf(nT) : 0.f;
        if (rS < io.sv_width) io.sv[rS] = nS > thr2 ? sqrtf(nS) : 0.f;
      }
      for (int k = Q + lg; k < io.sv_width; k += G) io.sv[k] = 0.f;
    }
    if (io.sweeps && lg == 0) *io.sweeps = sweeps;
  }
  // U = normalised columns
  const float iT = kT ? rsqrtf(nT) : 0.f;
  const float iS = kS ? rsqrtf(nS) : 0.f;
  const u64 iT2 = pack2(iT, iT), iS2 = pack2(iS, iS);
#pragma unroll
  for (int j = 0; j < NE; ++j) {
    T[j] = mul2(T[j], iT2);
    S[j] = mul2(S[j], iS2);
  }
  // A_n[(s*b + l), rank]
  if (valid) {
#pragma unroll
    for (int j = 0; j < NE; ++j) {
      float t0, t1, s0, s1;
      unpack2(T[j], t0, t1);
      unpack2(S[j], s0, s1);
      const int e = e0 + 2 * j, sj = e / B, la = e % B, lb = la + 1;
      if (la < b_rt) {
        if (rT < r_rt) io.A[(sj * b_rt + la) * r_rt + rT] = t0;
        if (rS < r_rt) io.A[(sj * b_rt + la) * r_rt + rS] = s0;
      }
      if (lb < b_rt) {
        if (rT < r_rt) io.A[(sj * b_rt + lb) * r_rt + rT] = t1;
        if (rS < r_rt) io.A[(sj * b_rt + lb) * r_rt + rS] = s1;
      }
    }
  }
  // Psi_{n+1}[r][c'] = sum_e U[e][r] M[e][c'],  M[(s,l)][c'] = Psi_n[l][s*Q + c']
  constexpr int Q2 = Q / 2;  // packed pairs per row of M
  u64 aT[Q2], aS[Q2];
#pragma unroll
  for (int q = 0; q < Q2; ++q) aT[q] = aS[q] = 0ull;
  {
#pragma unroll
    for (int j = 0; j < NE; ++j) {
      float t0, t1, s0, s1;
      unpack2(T[j], t0, t1);
      unpack2(S[j], s0, s1);
      const int e = e0 + 2 * j, sj = e / B, lj = e % B;
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        const u64 ut = p ? pack2(t1, t1) : pack2(t0, t0);
        const u64 us = p ? pack2(s1, s1) : pack2(s0, s0);
        const float* row = psi + (lj + p) * 2 * Q + sj * Q;
#pragma unroll
        for (int q4 = 0; q4 < Q / 4; ++q4) {
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
  if constexpr (LP == 2) {
    // sum the two element halves (lane pairs)
#pragma unroll
    for (int q = 0; q < Q2; ++q) {
      aT[q] = add2(aT[q], shfl2(aT[q], lane ^ 1));
      aS[q] = add2(aS[q], shfl2(aS[q], lane ^ 1));
    }
  }
  if (valid) {
    float* out = ws_out + (size_t)im * kWsStride;
    // LP == 2: lane sub = 0 writes row rT, sub = 1 row rS; LP == 1: the lane writes both
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      if (LP == 2 && w != sub) continue;
      const int r = w ? rS : rT;
      if (r < Q) {
#pragma unroll
        for (int q4 = 0; q4 < Q / 4; ++q4) {
          float4 v;
          const u64 a0 = w ? aS[2 * q4] : aT[2 * q4], a1 = w ? aS[2 * q4 + 1] : aT[2 * q4 + 1];
          unpack2(a0, v.x, v.y);
          unpack2(a1, v.z, v.w);
          *reinterpret_cast<float4*>(out + r * Q + 4 * q4) = v;
        }
      }
    }
  }
}

// ---- step 6, QR-preconditioned like step 5 -----------------------------------------
// M_6 (32 x 8): column-pivoted Gram-Schmidt in registers -> R (8 x 8, shared memory); one-sided Jacobi on
// the rows of R (8 vectors of 8 elements instead of 8 columns of 32); the converged vectors sigma_r v_r are
// the rows of Psi_7; U = M_6 V Sigma^-1 by one product (each lane owns the two vectors it iterated, so the
// coefficients are already in its registers) and one odd-even sweep of Gram-Schmidt projections.
// One lane per slot, 4 lanes per image, 8 images per warp, as hw_col_kernel<6>.
__global__ void __launch_bounds__(kHwWarps * 32, OC_HW_MINB)
hw_col6q_kernel(const __grid_constant__ EncodeParams P, const float* __restrict__ ws_in, float* __restrict__ ws_out) {
  constexpr int B = 16, Q = 8, NE = 16, IPW = 8;
  __shared__ __align__(16) float Rsm[kHwWarps * IPW * Q * Q];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int slot = lane & 3, grp = lane >> 2, gbase = lane & ~3;
  const int64_t img = ((int64_t)blockIdx.x * kHwWarps + warp) * IPW + grp;
  const bool valid = img < P.N;
  const int64_t im = valid ? img : 0;
  const float* psi = ws_in + (size_t)im * kWsStride;  // Psi_6 [16][16]
  float* R = Rsm + (size_t)(warp * IPW + grp) * Q * Q;  // R[step][column]
  // columns (2 slot, 2 slot + 1) of M_6; element e = s * B + l, packed pair j = (e = 2j, 2j + 1)
  u64 T[16], S[16];
  {
    const float* pc = psi + 2 * slot;
#pragma unroll
    for (int j = 0; j < NE; ++j) {
      const int e = 2 * j, sj = e / B, lj = e % B;
      float2 v0 = make_float2(0.f, 0.f), v1 = v0;
      if (valid) {
        v0 = __ldg(reinterpret_cast<const float2*>(pc + lj * 2 * Q + sj * Q));
        v1 = __ldg(reinterpret_cast<const float2*>(pc + (lj + 1) * 2 * Q + sj * Q));
      }
      T[j] = pack2(v0.x, v1.x);
      S[j] = pack2(v0.y, v1.y);
    }
  }
  float nT = dot16<0, NE>(T, T), nS = dot16<0, NE>(S, S);
  const float fro = group_sum<2>(nT + nS);
  const float thr2 = kFloor2 * fro;
  int nsteps = 0;  // R rows up to the last live pivot
#pragma unroll 1
  for (int step = 0; step < Q; ++step) {
    // pivot: largest remaining (cached, downdated) squared norm; ties -> lowest column
    unsigned key;
    {
      const unsigned kt = nT > 0.f ? ((__float_as_uint(nT) & ~15u) | (unsigned)(15 - 2 * slot)) : 0u;
      const unsigned ks = nS > 0.f ? ((__float_as_uint(nS) & ~15u) | (unsigned)(14 - 2 * slot)) : 0u;
      key = kt > ks ? kt : ks;
#pragma unroll
      for (int o = 1; o < 4; o <<= 1) {
        const unsigned other = __shfl_xor_sync(FULL, key, o);
        key = key > other ? key : other;
      }
    }
    const int pc = 15 - (int)(key & 15u);
    const int owner = gbase | ((pc >> 1) & 3);
    u64 qv[NE];
#pragma unroll
    for (int e = 0; e < NE; ++e) qv[e] = shfl2((pc & 1) ? S[e] : T[e], owner);
    float rT = dot16<0, NE>(qv, T), rS = dot16<0, NE>(qv, S);
    const float qn2 = __shfl_sync(FULL, (pc & 1) ? rS : rT, owner);  // the owner's own dot product is |q|^2
    const bool dead = key == 0u || !(qn2 > thr2);
    const float inv2 = dead ? 0.f : 1.f / qn2, invn = dead ? 0.f : rsqrtf(qn2);
    const bool isT = 2 * slot == pc, isS = 2 * slot + 1 == pc;
    if (!(nT > 0.f) || isT) rT = 0.f;  // finished columns (norm < 0) and the pivot itself stay
    if (!(nS > 0.f) || isS) rS = 0.f;
    const float cT = -rT * inv2, cS = -rS * inv2;
    const u64 cT2 = pack2(cT, cT), cS2 = pack2(cS, cS);
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      T[e] = fma2(cT2, qv[e], T[e]);
      S[e] = fma2(cS2, qv[e], S[e]);
    }
    {
      const float dg = qn2 * invn;
      *reinterpret_cast<float2*>(R + step * Q + 2 * slot) = make_float2(isT ? dg : rT * invn, isS ? dg : rS * invn);
    }
    nT = isT ? -1.f : (nT > 0.f ? fmaxf(fmaf(cT, rT, nT), 1e-37f) : nT);
    nS = isS ? -1.f : (nS > 0.f ? fmaxf(fmaf(cS, rS, nS), 1e-37f) : nS);
    if (!dead) nsteps = step + 1;
  }
  __syncwarp();
  // ---- Jacobi on the rows of R: this lane iterates rows 2 slot, 2 slot + 1 (8 elements each) ----
#pragma unroll
  for (int j = 0; j < 16; ++j) T[j] = S[j] = 0ull;
  {
    const float* r0 = R + (2 * slot) * Q;
    const float4 a0 = lds4(r0), a1 = lds4(r0 + 4), b0 = lds4(r0 + Q), b1 = lds4(r0 + Q + 4);
    T[0] = pack2(a0.x, a0.y); T[1] = pack2(a0.z, a0.w); T[2] = pack2(a1.x, a1.y); T[3] = pack2(a1.z, a1.w);
    S[0] = pack2(b0.x, b0.y); S[1] = pack2(b0.z, b0.w); S[2] = pack2(b1.x, b1.y); S[3] = pack2(b1.z, b1.w);
  }
  const int mlive = (nsteps + 1) >> 1;
  const int mlive_max = __reduce_max_sync(FULL, mlive);
  int sweeps = 0;
  nT = dot16<0, 4>(T, T);
  nS = dot16<0, 4>(S, S);
  if (mlive_max > 0) {
    sweeps = jacobi16<0, 4, 2>(T, S, nT, nS, thr2, lane, mlive, mlive_max);
    nT = dot16<0, 4>(T, T);
    nS = dot16<0, 4>(S, S);
  }
  int rT, rS;
  rank16<0, 2>(nT, nS, lane, rT, rS);
  const Out io = make_out(P, im, 6);
  const int r_rt = io.r_rt, b_rt = io.b_rt;
  const bool kT = nT > thr2 && rT < r_rt;
  const bool kS = nS > thr2 && rS < r_rt;
  if (valid) {
    if (io.sv) {
      if (rT < io.sv_width) io.sv[rT] = nT > thr2 ? sqrtf(nT) : 0.f;
      if (rS < io.sv_width) io.sv[rS] = nS > thr2 ? sqrtf(nS) : 0.f;
      for (int k = Q + slot; k < io.sv_width; k += 4) io.sv[k] = 0.f;
    }
    if (io.sweeps && slot == 0) *io.sweeps = sweeps;
    // Psi_7[rank][c] = sigma_r v_r[c]: the converged vector itself
    float* out = ws_out + (size_t)im * kWsStride;
    float4 t0, t1, s0, s1;
    unpack2(T[0], t0.x, t0.y); unpack2(T[1], t0.z, t0.w); unpack2(T[2], t1.x, t1.y); unpack2(T[3], t1.z, t1.w);
    unpack2(S[0], s0.x, s0.y); unpack2(S[1], s0.z, s0.w); unpack2(S[2], s1.x, s1.y); unpack2(S[3], s1.z, s1.w);
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(out + rT * Q) = kT ? t0 : z;
    *reinterpret_cast<float4*>(out + rT * Q + 4) = kT ? t1 : z;
    *reinterpret_cast<float4*>(out + rS * Q) = kS ? s0 : z;
    *reinterpret_cast<float4*>(out + rS * Q + 4) = kS ? s1 : z;
  }
  // ---- U = M_6 V Sigma^-1: coefficients (sigma v)[c] / sigma^2 of this lane's two vectors ----
  u64 cf[Q];  // cf[c] = (coefficient of vector T, of vector S)
  {
    const float iT = kT ? 1.f / nT : 0.f, iS = kS ? 1.f / nS : 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float a0, a1, b0, b1;
      unpack2(T[e], a0, a1);
      unpack2(S[e], b0, b1);
      cf[2 * e] = pack2(a0 * iT, b0 * iS);
      cf[2 * e + 1] = pack2(a1 * iT, b1 * iS);
    }
  }
  u64 UT[NE], US[NE];
#pragma unroll
  for (int j = 0; j < NE; ++j) {
    u64 acc[2] = {0ull, 0ull};  // (U_T, U_S) of elements e = 2j, 2j + 1
    const int e = 2 * j, sj = e / B, lj = e % B;
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const float* row = psi + (lj + p) * 2 * Q + sj * Q;
      float4 m0 = make_float4(0.f, 0.f, 0.f, 0.f), m1 = m0;
      if (valid) {
        m0 = ldg4(row);
        m1 = ldg4(row + 4);
      }
      acc[p] = fma2(cf[0], pack2(m0.x, m0.x), acc[p]);
      acc[p] = fma2(cf[1], pack2(m0.y, m0.y), acc[p]);
      acc[p] = fma2(cf[2], pack2(m0.z, m0.z), acc[p]);
      acc[p] = fma2(cf[3], pack2(m0.w, m0.w), acc[p]);
      acc[p] = fma2(cf[4], pack2(m1.x, m1.x), acc[p]);
      acc[p] = fma2(cf[5], pack2(m1.y, m1.y), acc[p]);
      acc[p] = fma2(cf[6], pack2(m1.z, m1.z), acc[p]);
      acc[p] = fma2(cf[7], pack2(m1.w, m1.w), acc[p]);
    }
    float a0, b0, a1, b1;
    unpack2(acc[0], a0, b0);
    unpack2(acc[1], a1, b1);
    UT[j] = pack2(a0, a1);
    US[j] = pack2(b0, b1);
  }
  // ---- polish (one odd-even sweep of projections), normalise, A_6[(s*b + l), rank] ----
  {
    float keyT = kT ? nT : 0.f, keyS = kS ? nS : 0.f;
    int tagT = rT, tagS = rS;
    polish_ring<NE>(UT, US, keyT, keyS, tagT, tagS, lane, 4, gbase, mlive, mlive_max);
    const bool tiny = (keyT > 0.f && keyT < 1e3f * thr2) || (keyS > 0.f && keyS < 1e3f * thr2);  // as in step 5
    if (__any_sync(FULL, tiny)) {
      polish_ring<NE>(UT, US, keyT, keyS, tagT, tagS, lane, 4, gbase, mlive, mlive_max);
      polish_ring<NE>(UT, US, keyT, keyS, tagT, tagS, lane, 4, gbase, mlive, mlive_max);
    }
    u64 qT = 0ull, qS = 0ull;
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      qT = fma2(UT[e], UT[e], qT);
      qS = fma2(US[e], US[e], qS);
    }
    const float jT = keyT > 0.f ? rsqrtf(hsum2(qT)) : 0.f;
    const float jS = keyS > 0.f ? rsqrtf(hsum2(qS)) : 0.f;
    if (valid) {
#pragma unroll
      for (int j = 0; j < NE; ++j) {
        float t0, t1, s0, s1;
        unpack2(UT[j], t0, t1);
        unpack2(US[j], s0, s1);
        const int e = 2 * j, sj = e / B, la = e % B, lb = la + 1;
        if (la < b_rt) {
          if (tagT < r_rt) io.A[(sj * b_rt + la) * r_rt + tagT] = t0 * jT;
          if (tagS < r_rt) io.A[(sj * b_rt + la) * r_rt + tagS] = s0 * jS;
        }
        if (lb < b_rt) {
          if (tagT < r_rt) io.A[(sj * b_rt + lb) * r_rt + tagT] = t1 * jT;
          if (tagS < r_rt) io.A[(sj * b_rt + lb) * r_rt + tagS] = s1 * jS;
        }
      }
    }
  }
}
inline int& qr_step6() {
  static int v = 1;
  return v;
}

// ---- step 5, QR-preconditioned ---------------------------------------------------
// M_5 (64 x 16, columns spread over the two lanes of a slot) = Q R by one pass of
// column-pivoted modified Gram-Schmidt in registers; the right singular vectors and
// the singular values come from one-sided Jacobi on the rows of the SMALL factor R
// (16 vectors of 16 elements, 5.1 sweeps; the 64-element columns took 7.0), whose
// converged vectors sigma_r v_r ARE the rows of Psi_6.  U = M_5 V Sigma^-1 is one
// product; its columns lose orthogonality like sigma_max / sigma_r, which one sweep
// of Gram-Schmidt projections (smaller sigma against larger, as in the row steps)
// restores (numpy model: 6e-4 -> 2e-7).  NE = packed pairs per lane and column:
// 13 when rows l >= 26 of Psi_5 are structurally zero (n_pix <= 832), else 16.
// PART 0: the whole step in one kernel.  PART 1 / (hw_col5b_kernel) / PART 3: the same cut in
// three - Gram-Schmidt | 16 x 16 Jacobi | U, polish, A_5 - so that the Jacobi, which is all
// rotation-parameter chain, runs with ONE lane per slot and four images per warp instead of
// two lanes per slot computing every chain twice.  R, then the converged vectors, their norms and
// ranks travel through the free part of the image's Psi_6 slot in the workspace:
//   floats 0..255 Psi_6 | 256..511 R / converged vectors | 512 thr2 | 513 nsteps | 514 sweeps |
//   520..535 squared norms | 536..551 ranks
constexpr int kC5R = 256, kC5Meta = 512, kC5N2 = 520, kC5Rank = 536;
template <int NE, int PART = 0>
__global__ void __launch_bounds__(kHwWarps * 32, 4)
hw_col5q_kernel(const __grid_constant__ EncodeParams P, const float* __restrict__ ws_in, float* __restrict__ ws_out) {
  constexpr int Q = 16;
  __shared__ __align__(16) float Rsm[PART == 0 ? kHwWarps * 2 * 256 : 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lg = lane & 15, hb = lane & 16, h = lane >> 4;
  const int64_t img = ((int64_t)blockIdx.x * kHwWarps + warp) * 2 + h;
  const bool valid = img < P.N;
  const int64_t im = valid ? img : 0;
  const float* psi = ws_in + (size_t)im * kWsStride;  // [32][2Q]
  const int slot = lg >> 1, sub = lg & 1;
  float* slot6 = ws_out + (size_t)im * kWsStride;  // Psi_6 slot of the image (PART != 0: R and meta too)
  float* R = PART == 0 ? Rsm + (size_t)(warp * 2 + h) * 256 : slot6 + kC5R;  // R[step][column]
  if constexpr (PART == 3) {
    // M_5 is first needed after a dependent round trip for the Jacobi results: start its
    // 2 NE rows of 128 B towards L2 now (lane lg takes rows 2 lg, 2 lg + 1)
    if (valid && lg < NE) {
      asm volatile("prefetch.global.L1 [%0];" ::"l"(psi + (2 * lg) * 2 * Q));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(psi + (2 * lg + 1) * 2 * Q));
    }
  }
  // lane's elements: (s = sub, l = 0 .. 2 NE - 1); packed pair j = (l = 2j, 2j + 1)
  u64 T[16], S[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) T[j] = S[j] = 0ull;
  float nT = 0.f, nS = 0.f, thr2 = 0.f;
  int nsteps = 0;  // R rows up to the last live pivot
  if constexpr (PART != 3) {
  {
    const float* pc = psi + sub * Q + 2 * slot;
#pragma unroll
    for (int j = 0; j < NE; ++j) {
      float2 v0 = make_float2(0.f, 0.f), v1 = v0;
      if (valid) {
        v0 = __ldg(reinterpret_cast<const float2*>(pc + (2 * j) * 2 * Q));
        v1 = __ldg(reinterpret_cast<const float2*>(pc + (2 * j + 1) * 2 * Q));
      }
      T[j] = pack2(v0.x, v1.x);
      S[j] = pack2(v0.y, v1.y);
    }
  }
  nT = dot16<1, NE>(T, T);
  nS = dot16<1, NE>(S, S);
  const float fro = group_sum<4>(sub == 0 ? nT + nS : 0.f);
  thr2 = kFloor2 * fro;
  // ---- column-pivoted MGS; R[step][c] -> shared (PART 1: workspace) ----
  {
#pragma unroll 1
    for (int step = 0; step < Q; ++step) {
      // pivot: largest remaining (cached, downdated) squared norm; ties -> lowest column
      unsigned key = 0u;
      {
        const unsigned kt = nT > 0.f ? ((__float_as_uint(nT) & ~15u) | (unsigned)(15 - 2 * slot)) : 0u;
        const unsigned ks = nS > 0.f ? ((__float_as_uint(nS) & ~15u) | (unsigned)(14 - 2 * slot)) : 0u;
        key = kt > ks ? kt : ks;
#pragma unroll
        for (int o = 2; o < 16; o <<= 1) {
          const unsigned other = __shfl_xor_sync(FULL, key, o);
          key = key > other ? key : other;
        }
      }
      const int pc = 15 - (int)(key & 15u);
      const int owner = hb | (((pc >> 1) << 1) | sub);
      u64 qv[NE];
#pragma unroll
      for (int e = 0; e < NE; ++e) qv[e] = shfl2((pc & 1) ? S[e] : T[e], owner);
      float rT = dot16<1, NE>(qv, T), rS = dot16<1, NE>(qv, S);
      // |q|^2 is the owner's own dot product (its column IS q): one shuffle instead of a third dot.
      // The cached norms only ORDER the pivots; a column whose true residual is below the floor is
      // retired as dead (zero row of R) and the factorisation goes on with the others
      const float qn2 = __shfl_sync(FULL, (pc & 1) ? rS : rT, owner);
      const bool dead = key == 0u || !(qn2 > thr2);
      const float inv2 = dead ? 0.f : 1.f / qn2, invn = dead ? 0.f : rsqrtf(qn2);
      const bool isT = 2 * slot == pc, isS = 2 * slot + 1 == pc;
      if (!(nT > 0.f) || isT) rT = 0.f;  // finished columns (norm < 0) and the pivot itself stay
      if (!(nS > 0.f) || isS) rS = 0.f;
      const float cT = -rT * inv2, cS = -rS * inv2;
      const u64 cT2 = pack2(cT, cT), cS2 = pack2(cS, cS);
#pragma unroll
      for (int e = 0; e < NE; ++e) {
        T[e] = fma2(cT2, qv[e], T[e]);
        S[e] = fma2(cS2, qv[e], S[e]);
      }
      if (sub == 0 && (PART == 0 || valid)) {
        const float dg = qn2 * invn;
        *reinterpret_cast<float2*>(R + step * Q + 2 * slot) =
            make_float2(isT ? dg : rT * invn, isS ? dg : rS * invn);
      }
      // downdate the cached norms; the pivot column is finished
      nT = isT ? -1.f : (nT > 0.f ? fmaxf(fmaf(cT, rT, nT), 1e-37f) : nT);
      nS = isS ? -1.f : (nS > 0.f ? fmaxf(fmaf(cS, rS, nS), 1e-37f) : nS);
      if (!dead) nsteps = step + 1;
    }
  }
  }  // PART != 3
  if constexpr (PART == 1) {
    if (valid && lg == 0) {
      slot6[kC5Meta] = thr2;
      reinterpret_cast<int*>(slot6)[kC5Meta + 1] = nsteps;
    }
    return;
  }
  __syncwarp();
  // ---- Jacobi on the rows of R (vector j = R[j][.], 8 elements per lane) ----
#pragma unroll
  for (int j = 0; j < 16; ++j) T[j] = S[j] = 0ull;
  if (PART == 0 || valid) {
    const float* r0 = R + (2 * slot) * Q + 8 * sub;  // PART 3: the converged vectors hw_col5b_kernel left here
    const float4 a0 = ld4<PART != 0>(r0), a1 = ld4<PART != 0>(r0 + 4), b0 = ld4<PART != 0>(r0 + Q),
                 b1 = ld4<PART != 0>(r0 + Q + 4);
    T[0] = pack2(a0.x, a0.y); T[1] = pack2(a0.z, a0.w); T[2] = pack2(a1.x, a1.y); T[3] = pack2(a1.z, a1.w);
    S[0] = pack2(b0.x, b0.y); S[1] = pack2(b0.z, b0.w); S[2] = pack2(b1.x, b1.y); S[3] = pack2(b1.z, b1.w);
  }
  int sweeps = 0, rT = 0, rS = 0;
  if constexpr (PART == 3) {
    thr2 = valid ? slot6[kC5Meta] : 0.f;
    nsteps = valid ? reinterpret_cast<const int*>(slot6)[kC5Meta + 1] : 0;
    nT = valid ? slot6[kC5N2 + 2 * slot] : 0.f;
    nS = valid ? slot6[kC5N2 + 2 * slot + 1] : 0.f;
    rT = valid ? reinterpret_cast<const int*>(slot6)[kC5Rank + 2 * slot] : 2 * slot;
    rS = valid ? reinterpret_cast<const int*>(slot6)[kC5Rank + 2 * slot + 1] : 2 * slot + 1;
  }
  const int mlive = (nsteps + 1) >> 1;
  int mlive_max = mlive;
  mlive_max = max(mlive_max, __shfl_xor_sync(FULL, mlive_max, 16));
  if constexpr (PART == 0) {
    nT = dot16<1, 4>(T, T);
    nS = dot16<1, 4>(S, S);
    if (mlive_max > 0) {
      sweeps = jacobi16<1, 4, 4>(T, S, nT, nS, thr2, lane, mlive, mlive_max);
      nT = dot16<1, 4>(T, T);
      nS = dot16<1, 4>(S, S);
    }
    rank16<1, 4>(nT, nS, lane, rT, rS);
  }
  const Out io = make_out(P, im, 5);
  const int r_rt = io.r_rt, b_rt = io.b_rt;
  const bool kT = nT > thr2 && rT < r_rt;
  const bool kS = nS > thr2 && rS < r_rt;
  if (PART == 0 && valid) {
    if (io.sv) {
      if (sub == 0) {
        if (rT < io.sv_width) io.sv[rT] = nT > thr2 ? sqrtf(nT) : 0.f;
        if (rS < io.sv_width) io.sv[rS] = nS > thr2 ? sqrtf(nS) : 0.f;
      }
      for (int k = Q + lg; k < io.sv_width; k += 16) io.sv[k] = 0.f;
    }
    if (io.sweeps && lg == 0) *io.sweeps = sweeps;
    // Psi_6[rank][c] = sigma_r v_r[c]: the converged vector itself
    float* out = ws_out + (size_t)im * kWsStride + 8 * sub;
    float4 t0, t1, s0, s1;
    unpack2(T[0], t0.x, t0.y); unpack2(T[1], t0.z, t0.w); unpack2(T[2], t1.x, t1.y); unpack2(T[3], t1.z, t1.w);
    unpack2(S[0], s0.x, s0.y); unpack2(S[1], s0.z, s0.w); unpack2(S[2], s1.x, s1.y); unpack2(S[3], s1.z, s1.w);
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rT < Q) {
      *reinterpret_cast<float4*>(out + rT * Q) = kT ? t0 : z;
      *reinterpret_cast<float4*>(out + rT * Q + 4) = kT ? t1 : z;
    }
    if (rS < Q) {
      *reinterpret_cast<float4*>(out + rS * Q) = kS ? s0 : z;
      *reinterpret_cast<float4*>(out + rS * Q + 4) = kS ? s1 : z;
    }
  }
  // ---- U = M_5 V Sigma^-1: coefficients (sigma v)[c] / sigma^2 of both columns, all 16 c ----
  u64 cf[16];  // cf[c] = (coefficient of column T, of column S)
  {
    const float iT = kT ? 1.f / nT : 0.f, iS = kS ? 1.f / nS : 0.f;
    float vt[8], vs[8];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      unpack2(T[e], vt[2 * e], vt[2 * e + 1]);
      unpack2(S[e], vs[2 * e], vs[2 * e + 1]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float mt = vt[i] * iT, ms = vs[i] * iS;
      const float ot = __shfl_xor_sync(FULL, mt, 1), os = __shfl_xor_sync(FULL, ms, 1);
      // this lane holds c = 8 sub + i, its partner c = 8 (1 - sub) + i
      cf[i] = sub ? pack2(ot, os) : pack2(mt, ms);
      cf[8 + i] = sub ? pack2(mt, ms) : pack2(ot, os);
    }
  }
  u64 UT[NE], US[NE];
  {
    const float* pr = psi + sub * Q;
#pragma unroll
    for (int j = 0; j < NE; ++j) {
      u64 acc[2] = {0ull, 0ull};  // (U_T, U_S) of elements l = 2j, 2j + 1
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        const float* row = pr + (2 * j + p) * 2 * Q;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
          if (valid) m = ldg4(row + 4 * q4);
          acc[p] = fma2(cf[4 * q4], pack2(m.x, m.x), acc[p]);
          acc[p] = fma2(cf[4 * q4 + 1], pack2(m.y, m.y), acc[p]);
          acc[p] = fma2(cf[4 * q4 + 2], pack2(m.z, m.z), acc[p]);
          acc[p] = fma2(cf[4 * q4 + 3], pack2(m.w, m.w), acc[p]);
        }
      }
      float a0, b0, a1, b1;
      unpack2(acc[0], a0, b0);
      unpack2(acc[1], a1, b1);
      UT[j] = pack2(a0, a1);
      US[j] = pack2(b0, b1);
    }
  }
  // ---- polish (one odd-even sweep of projections), normalise, A_5[(s*b + l), rank] ----
  {
    float keyT = kT ? nT : 0.f, keyS = kS ? nS : 0.f;
    int tagT = rT, tagS = rS;
    polish16<1, NE, true>(UT, US, keyT, keyS, tagT, tagS, lane, mlive, mlive_max);
    // columns with sigma < 3e-5 sigma_max start ~1e-2 off, and what one pass leaves between two of
    // them is second order in that (1e-3 measured on low-contrast images): two more passes, only for
    // warps that have such columns (none on MNIST-like data)
    const bool tiny = (keyT > 0.f && keyT < 1e3f * thr2) || (keyS > 0.f && keyS < 1e3f * thr2);  // sigma^2 < 1e-9 |M|_F^2
    if (__any_sync(FULL, tiny)) {
      polish16<1, NE, true>(UT, US, keyT, keyS, tagT, tagS, lane, mlive, mlive_max);
      polish16<1, NE, true>(UT, US, keyT, keyS, tagT, tagS, lane, mlive, mlive_max);
    }
    u64 qT = 0ull, qS = 0ull;
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      qT = fma2(UT[e], UT[e], qT);
      qS = fma2(US[e], US[e], qS);
    }
    float sT = hsum2(qT), sS = hsum2(qS);
    sT += __shfl_xor_sync(FULL, sT, 1);
    sS += __shfl_xor_sync(FULL, sS, 1);
    const float jT = keyT > 0.f ? rsqrtf(sT) : 0.f;
    const float jS = keyS > 0.f ? rsqrtf(sS) : 0.f;
    if (valid && b_rt == 32 && r_rt == Q && (reinterpret_cast<uintptr_t>(io.A) & 15) == 0) {
      // D >= 32 (the untruncated site, 2 x 32 x 16): no bounds to test, the structurally zero rows
      // l >= 2 NE of both s as 128-bit stores (the run-time-shaped code below was 28 % of this kernel's samples)
      float* A = io.A + sub * 32 * Q;
#pragma unroll
      for (int j = 0; j < NE; ++j) {
        float t0, t1, s0, s1;
        unpack2(UT[j], t0, t1);
        unpack2(US[j], s0, s1);
        A[(2 * j) * Q + tagT] = t0 * jT;
        A[(2 * j) * Q + tagS] = s0 * jS;
        A[(2 * j + 1) * Q + tagT] = t1 * jT;
        A[(2 * j + 1) * Q + tagS] = s1 * jS;
      }
      constexpr int Z4 = (32 - 2 * NE) * Q / 4;  // float4 per s
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      if constexpr (Z4 > 0) {
#pragma unroll
        for (int q = lg; q < 2 * Z4; q += 16)
          *reinterpret_cast<float4*>(io.A + ((q / Z4) * 32 + 2 * NE) * Q + 4 * (q % Z4)) = z;
      }
    } else if (valid) {
#pragma unroll
      for (int j = 0; j < NE; ++j) {
        float t0, t1, s0, s1;
        unpack2(UT[j], t0, t1);
        unpack2(US[j], s0, s1);
        const int la = 2 * j, lb = la + 1;
        if (la < b_rt) {
          if (tagT < r_rt) io.A[(sub * b_rt + la) * r_rt + tagT] = t0 * jT;
          if (tagS < r_rt) io.A[(sub * b_rt + la) * r_rt + tagS] = s0 * jS;
        }
        if (lb < b_rt) {
          if (tagT < r_rt) io.A[(sub * b_rt + lb) * r_rt + tagT] = t1 * jT;
          if (tagS < r_rt) io.A[(sub * b_rt + lb) * r_rt + tagS] = s1 * jS;
        }
      }
      // rows l >= 2 NE are structurally zero
      const int nz = (b_rt - 2 * NE) * r_rt;
      if (nz > 0)
        for (int e = lg; e < 2 * nz; e += 16) io.A[((e / nz) * b_rt + 2 * NE) * r_rt + e % nz] = 0.f;
    }
  }
}

// ---- step 5, middle part: Jacobi on the rows of R with one lane per slot, 4 images per warp ----
__global__ void __launch_bounds__(kHwWarps * 32, 6)
hw_col5b_kernel(const __grid_constant__ EncodeParams P, float* __restrict__ ws_out) {
  constexpr int Q = 16;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sl = lane & 7;
  const int64_t img = ((int64_t)blockIdx.x * kHwWarps + warp) * 4 + (lane >> 3);
  const bool valid = img < P.N;
  const int64_t im = valid ? img : 0;
  float* slot6 = ws_out + (size_t)im * kWsStride;
  float* r0 = slot6 + kC5R + (2 * sl) * Q;
  u64 A[8], Bv[8];
#pragma unroll
  for (int q4 = 0; q4 < 4; ++q4) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), bq = a;
    if (valid) {
      a = *reinterpret_cast<const float4*>(r0 + 4 * q4);
      bq = *reinterpret_cast<const float4*>(r0 + Q + 4 * q4);
    }
    A[2 * q4] = pack2(a.x, a.y); A[2 * q4 + 1] = pack2(a.z, a.w);
    Bv[2 * q4] = pack2(bq.x, bq.y); Bv[2 * q4 + 1] = pack2(bq.z, bq.w);
  }
  const float thr2 = valid ? slot6[kC5Meta] : 0.f;
  const int nsteps = valid ? reinterpret_cast<const int*>(slot6)[kC5Meta + 1] : 0;
  float nA = dot16<0, 8>(A, A), nB = dot16<0, 8>(Bv, Bv);
  const int ml = (nsteps + 1) >> 1;
  const int mlmax = __reduce_max_sync(FULL, ml);
  int sw = 0;
  if (mlmax > 0) {
    sw = jacobi_ring<8>(A, Bv, nA, nB, thr2, lane, 8, lane & ~7, ml, mlmax);
    nA = dot16<0, 8>(A, A);
    nB = dot16<0, 8>(Bv, Bv);
  }
  int rA, rB;
  rank16<0, 3>(nA, nB, lane, rA, rB);
  if (!valid) return;
  const Out io = make_out(P, im, 5);
  const bool kA = nA > thr2 && rA < io.r_rt, kB = nB > thr2 && rB < io.r_rt;
  if (io.sv) {
    if (rA < io.sv_width) io.sv[rA] = nA > thr2 ? sqrtf(nA) : 0.f;
    if (rB < io.sv_width) io.sv[rB] = nB > thr2 ? sqrtf(nB) : 0.f;
    for (int k = Q + sl; k < io.sv_width; k += 8) io.sv[k] = 0.f;
  }
  if (io.sweeps && sl == 0) *io.sweeps = sw;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int q4 = 0; q4 < 4; ++q4) {
    float4 a, bq;
    unpack2(A[2 * q4], a.x, a.y); unpack2(A[2 * q4 + 1], a.z, a.w);
    unpack2(Bv[2 * q4], bq.x, bq.y); unpack2(Bv[2 * q4 + 1], bq.z, bq.w);
    // the converged vectors for the third part, and as rows of Psi_6 (sigma_r v_r)
    *reinterpret_cast<float4*>(r0 + 4 * q4) = a;
    *reinterpret_cast<float4*>(r0 + Q + 4 * q4) = bq;
    if (rA < Q) *reinterpret_cast<float4*>(slot6 + rA * Q + 4 * q4) = kA ? a : z;
    if (rB < Q) *reinterpret_cast<float4*>(slot6 + rB * Q + 4 * q4) = kB ? bq : z;
  }
  slot6[kC5N2 + 2 * sl] = nA;
  slot6[kC5N2 + 2 * sl + 1] = nB;
  reinterpret_cast<int*>(slot6)[kC5Rank + 2 * sl] = rA;
  reinterpret_cast<int*>(slot6)[kC5Rank + 2 * sl + 1] = rB;
}

// ---- steps 8 and 9, one THREAD per image: Psi_8 is 4 x 4 ----
__global__ void __launch_bounds__(128) hw_tail89_kernel(const __grid_constant__ EncodeParams P,
                                                        const float* __restrict__ ws_in) {
  const int64_t img = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (img >= P.N) return;
  const float* psi = ws_in + (size_t)img * kWsStride;  // [4][2*2]
  float m[16];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 v = ldg4(psi + 4 * i);
    m[4 * i] = v.x; m[4 * i + 1] = v.y; m[4 * i + 2] = v.z; m[4 * i + 3] = v.w;
  }
  // step 8: M_8[(s,l)][c] = Psi_8[l][2s + c]: two columns of length 8; one rotation orthogonalises them
  float a[8], b[8];
#pragma unroll
  for (int s = 0; s < 2; ++s)
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      a[s * 4 + l] = m[l * 4 + 2 * s];
      b[s * 4 + l] = m[l * 4 + 2 * s + 1];
    }
  float aa = 0.f, bb = 0.f, ab = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    aa = fmaf(a[e], a[e], aa);
    bb = fmaf(b[e], b[e], bb);
    ab = fmaf(a[e], b[e], ab);
  }
  const float thr2 = kFloor2 * (aa + bb);
  int sweeps8 = 0;
  if (aa + bb > 0.f) {
    // two columns: one rotation orthogonalises them in exact arithmetic; when their norms are
    // graded (|a| / |b| ~ 1e-4 for low-contrast images) the fp32 rotation leaves cos ~ 1e-4, so
    // rotate again until cos^2 <= 1e-12 (at most 3 times)
#pragma unroll 1
    for (int it = 0; it < 3; ++it) {
      if (!(ab * ab > kTol2 * aa * bb && aa > thr2 && bb > thr2)) break;
      ++sweeps8;
      float c, s;
      jacobi_cs(aa, bb, ab, c, s);
      aa = bb = ab = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float x = a[e], y = b[e];
        a[e] = c * x - s * y;
        b[e] = s * x + c * y;
        aa = fmaf(a[e], a[e], aa);
        bb = fmaf(b[e], b[e], bb);
        ab = fmaf(a[e], b[e], ab);
      }
    }
    if (sweeps8 == 0) sweeps8 = 1;
  }
  {
    const Out io = make_out(P, img, 8);
    const bool afirst = aa >= bb;  // stable descending order
    const float n0 = afirst ? aa : bb, n1 = afirst ? bb : aa;
    const bool k0 = n0 > thr2 && 0 < io.r_rt, k1 = n1 > thr2 && 1 < io.r_rt;
    const float i0 = k0 ? rsqrtf(n0) : 0.f, i1 = k1 ? rsqrtf(n1) : 0.f;
    float u0[8], u1[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      u0[e] = (afirst ? a[e] : b[e]) * i0;
      u1[e] = (afirst ? b[e] : a[e]) * i1;
    }
    if (io.sv) {
      for (int k = 0; k < io.sv_width; ++k) io.sv[k] = 0.f;
      io.sv[0] = n0 > thr2 ? sqrtf(n0) : 0.f;
      if (io.sv_width > 1) io.sv[1] = n1 > thr2 ? sqrtf(n1) : 0.f;
    }
    if (io.sweeps) *io.sweeps = sweeps8;
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
      for (int l = 0; l < 4; ++l)
        if (l < io.b_rt) {
          if (0 < io.r_rt) io.A[(s * io.b_rt + l) * io.r_rt + 0] = u0[s * 4 + l];
          if (1 < io.r_rt) io.A[(s * io.b_rt + l) * io.r_rt + 1] = u1[s * 4 + l];
        }
    // Psi_9[r][c] = sum_e U[e][r] M_8[e][c]
    float p[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        const float mc0 = m[l * 4 + 2 * s], mc1 = m[l * 4 + 2 * s + 1];
        p[0] = fmaf(u0[s * 4 + l], mc0, p[0]);
        p[1] = fmaf(u0[s * 4 + l], mc1, p[1]);
        p[2] = fmaf(u1[s * 4 + l], mc0, p[2]);
        p[3] = fmaf(u1[s * 4 + l], mc1, p[3]);
      }
    // step 9: M_9[(s,l)] = Psi_9[l][s] (4 x 1) -> A_9 = M / |M|; the 1 x 1 S.V is dropped (unit norm)
    const Out io9 = make_out(P, img, 9);
    const float nn = p[0] * p[0] + p[1] * p[1] + p[2] * p[2] + p[3] * p[3];
    const float inv = nn > 0.f ? rsqrtf(nn) : 0.f;
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
      for (int l = 0; l < 2; ++l)
        if (l < io9.b_rt) io9.A[s * io9.b_rt + l] = p[l * 2 + s] * inv;
    if (io9.sv) {
      for (int k = 0; k < io9.sv_width; ++k) io9.sv[k] = 0.f;
      io9.sv[0] = sqrtf(nn);
    }
    if (io9.sweeps) *io9.sweeps = 0;
  }
}

// D >= 8: bonds 1..3 are the natural 2, 4, 8, so steps 0..2 never truncate (the Gram front end
// needs that); smaller D runs the one-image-per-warp kernels of oc_encode_fast.cuh
inline bool supported(const EncodeParams& P) { return P.L == 10 && P.D >= 8; }

}  // namespace hw
namespace gram {  // oc_encode_gram.cuh: steps 0..3 from the 16 x 16 Gram matrix
inline int& enabled();
inline cudaError_t launch_head(const EncodeParams& P, double* gws, float* ws_psi4, uint8_t* keys, cudaStream_t st);
}  // namespace gram
namespace hw {

// steps 0..5 with the half-warp kernels, 6..9 with the v3 tail
inline int& sort_pairs() {
  static int v = 1;
  return v;
}
inline int& pack_step4() {
  static int v = 1;
  return v;
}
inline int& mix_step4() {
  static int v = 1;
  return v;
}
inline int& mgs_step4() {
  static int v = 1;
  return v;
}
inline int& qr_step5() {
  static int v = 1;
  return v;
}
inline int& split_step5() {
  static int v = 1;
  return v;
}
inline cudaError_t launch_chain(const EncodeParams& P, float* ws0, float* ws1, void* sort_ws, int64_t chunk,
                                int sms, cudaStream_t st) {
  const int per_cta = kHwWarps * 2;
  const int grid = (int)((P.N + per_cta - 1) / per_cta);
  const size_t smem_head = (size_t)per_cta * (1024 + ysm_floats(2)) * sizeof(float);
  const size_t smem3 = (size_t)per_cta * ysm_floats(3) * sizeof(float);
  const size_t smem4 = (size_t)per_cta * ysm_floats(4) * sizeof(float);
  uint8_t* keys = reinterpret_cast<uint8_t*>(sort_ws);
  int32_t* perm = reinterpret_cast<int32_t*>(keys + ((chunk + 255) & ~(int64_t)255));
  int32_t* bins = perm + encode_perm_ints(chunk);
  const bool sorted = sort_pairs() != 0;
  EncodeEvents& ee = encode_events();
  ee.mark(0, st);
  if (gram::enabled()) {
    // steps 0..3 from the Gram matrix; G_3 (2 KB / image) borrows ws0, which step 4 overwrites later
    const cudaError_t e = gram::launch_head(P, reinterpret_cast<double*>(ws0), ws1, sorted ? keys : nullptr, st);
    if (e != cudaSuccess) return e;
    ee.mark(1, st);
  } else {
    hw_head_kernel<<<grid, kHwWarps * 32, smem_head, st>>>(P, ws0);
    ee.mark(1, st);
    if (sorted)
      hw_row_kernel<3, false, true><<<grid, kHwWarps * 32, smem3, st>>>(P, ws0, ws1, nullptr, nullptr, keys);
    else
      hw_row_kernel<3, false, false><<<grid, kHwWarps * 32, smem3, st>>>(P, ws0, ws1, nullptr, nullptr, nullptr);
  }
  // mark 2 (start of the step-4 kernel) is set behind the sort launches below
  if (sorted && pack_step4()) {
    // step 4 costs (live rows) x (sweeps) per WARP: floor(32 / k) images of equal k = live rows / 2 per warp
    const int64_t max_warps = P.N / 2 + kSortBins + 1;
    cudaMemsetAsync(bins, 0, (size_t)4 * kSortBins * sizeof(int32_t), st);
    cudaMemsetAsync(perm, 0xff, (size_t)4 * max_warps * sizeof(int32_t), st);
    const int tb = 256, gb = (int)((P.N + tb - 1) / tb);
    hw_sort_hist_kernel<<<gb, tb, 0, st>>>(keys, P.N, bins);
    hw_pack_prefix_kernel<<<1, 32, 0, st>>>(bins, mgs_step4() && mix_step4());
    hw_pack_scatter_kernel<<<gb, tb, 0, st>>>(keys, P.N, bins, perm);
    const int grid4 = (int)((max_warps + kHwWarps - 1) / kHwWarps);
    const size_t smem4p = (size_t)kHwWarps * 64 * 34 * sizeof(float);
    ee.mark(2, st);
    if (mgs_step4())
      hw_row4q_kernel<<<grid4, kHwWarps * 32, (size_t)kHwWarps * 64 * kQ4RS * sizeof(float), st>>>(P, ws1, ws0, perm,
                                                                                                bins, keys);
    else
      hw_row4p_kernel<<<grid4, kHwWarps * 32, smem4p, st>>>(P, ws1, ws0, perm, bins, keys);
  } else if (sorted) {
    // step 4 costs (live rows) x (sweeps) per WARP: pair images with equal numbers of live rows
    cudaMemsetAsync(bins, 0, (size_t)4 * kSortBins * sizeof(int32_t), st);
    const int tb = 256, gb = (int)((P.N + tb - 1) / tb);
    hw_sort_hist_kernel<<<gb, tb, 0, st>>>(keys, P.N, bins);
    hw_sort_prefix_kernel<<<1, 32, 0, st>>>(bins, perm);
    hw_sort_scatter_kernel<<<gb, tb, 0, st>>>(keys, P.N, bins, perm);
    const int grid4 = (int)((P.N + kSortBins + per_cta - 1) / per_cta);
    ee.mark(2, st);
    hw_row_kernel<4, true, false><<<grid4, kHwWarps * 32, smem4, st>>>(P, ws1, ws0, perm, bins + 3 * kSortBins, nullptr);
  } else {
    ee.mark(2, st);
    hw_row_kernel<4, false, false><<<grid, kHwWarps * 32, smem4, st>>>(P, ws1, ws0, nullptr, nullptr, nullptr);
  }
  ee.mark(3, st);
  if (qr_step5() && split_step5()) {
    // Gram-Schmidt | 16 x 16 Jacobi (4 images per warp) | U, polish, A_5
    const int gridb = (int)((P.N + 4 * kHwWarps - 1) / (4 * kHwWarps));
    if (P.n_pix <= 832) hw_col5q_kernel<13, 1><<<grid, kHwWarps * 32, 0, st>>>(P, ws0, ws1);
    else hw_col5q_kernel<16, 1><<<grid, kHwWarps * 32, 0, st>>>(P, ws0, ws1);
    hw_col5b_kernel<<<gridb, kHwWarps * 32, 0, st>>>(P, ws1);
    if (P.n_pix <= 832) hw_col5q_kernel<13, 3><<<grid, kHwWarps * 32, 0, st>>>(P, ws0, ws1);
    else hw_col5q_kernel<16, 3><<<grid, kHwWarps * 32, 0, st>>>(P, ws0, ws1);
  } else if (qr_step5() && P.n_pix <= 832)
    hw_col5q_kernel<13><<<grid, kHwWarps * 32, 0, st>>>(P, ws0, ws1);
  else if (qr_step5())
    hw_col5q_kernel<16><<<grid, kHwWarps * 32, 0, st>>>(P, ws0, ws1);
  else
    hw_col_kernel<5><<<grid, kHwWarps * 32, 0, st>>>(P, ws0, ws1);
  ee.mark(4, st);
  // tail: steps 6 and 7 with 8 / 16 images per warp, steps 8-9 one thread per image
  const int ipc6 = kHwWarps * ColShape<6>::IPW, ipc7 = kHwWarps * ColShape<7>::IPW;
  if (qr_step6()) hw_col6q_kernel<<<(int)((P.N + ipc6 - 1) / ipc6), kHwWarps * 32, 0, st>>>(P, ws1, ws0);
  else hw_col_kernel<6><<<(int)((P.N + ipc6 - 1) / ipc6), kHwWarps * 32, 0, st>>>(P, ws1, ws0);
  hw_col_kernel<7><<<(int)((P.N + ipc7 - 1) / ipc7), kHwWarps * 32, 0, st>>>(P, ws0, ws1);
  hw_tail89_kernel<<<(int)((P.N + 127) / 128), 128, 0, st>>>(P, ws1);
  ee.mark(5, st);
  (void)sms;
  return cudaGetLastError();
}

}  // namespace hw
}  // namespace oc
