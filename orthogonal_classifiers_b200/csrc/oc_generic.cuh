// Generic (runtime-shaped) kernels: any L <= 10, any D, any one-hairy-site
// classifier.  One warp per image, everything staged in shared memory.
// These are the correctness baseline and the fallback for shapes the
// specialised kernels (oc_encode_fast.cuh / oc_overlap_fast.cuh) do not cover.
//
// Encoder algorithm (tools.py:217-221 + xmps left_from_state / truncation):
//   Psi_0 = padded image (1 x 2^L).  For n = 0..L-1:
//     M_n[(s,l), c] = Psi_n[l, s*q + c]           (p = 2 b_n rows, q = 2^(L-n-1) cols)
//     one-sided (Hestenes) Jacobi SVD of M_n:
//       p <= q  "row mode": rotate the rows; converged rows ARE S.V^T (next
//                Psi) and the accumulated rotations are U (= A_n);
//       p >  q  "col mode": rotate the columns; converged columns are U.S,
//                normalised -> U (= A_n); next Psi = U^T M_n.
//     sort by singular value, keep b_{n+1} = min(p, q, D).
//   The final 1x1 S.V is dropped (unit-norm MPS).
#pragma once
#include "oc_b200.h"
#include "oc_common.cuh"

namespace oc {

struct EncodeParams {
  const void* images;
  int in_dtype;
  int64_t N;
  int n_pix;
  int64_t ld;
  int L;
  int D;  // already clamped to [1, 2^L]
  void* mps_out;
  void* svals_out;  // nullable
  int32_t* sweeps_out;  // nullable
  int64_t mps_stride;   // elements per image
  int64_t site_off[OC_MAX_L + 1];
  int bonds[OC_MAX_L + 1];
  int sv_width;
};

template <typename T>
__global__ void __launch_bounds__(128) encode_generic_kernel(const __grid_constant__ EncodeParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int warps_per_block = blockDim.x >> 5;
  const int cap = 1 << P.L;  // elements per buffer
  T* base = reinterpret_cast<T*>(smem_raw) + (size_t)warp * (3 * cap + 64);
  T* bufP = base;             // Psi_n
  T* bufQ = base + cap;       // work / output
  T* bufU = base + 2 * cap;   // U accumulation / output
  T* sig2 = base + 3 * cap;   // NV norms^2 (<= 32) + 32 spare

  for (int64_t img = (int64_t)blockIdx.x * warps_per_block + warp; img < P.N;
       img += (int64_t)gridDim.x * warps_per_block) {
    const char* src = reinterpret_cast<const char*>(P.images);
    const int esz = P.in_dtype == 0 ? 4 : (P.in_dtype == 1 ? 8 : 1);
    const void* row = src + (size_t)img * P.ld * esz;
    for (int i = lane; i < cap; i += 32)
      bufP[i] = i < P.n_pix ? load_pixel<T>(row, P.in_dtype, i) : T(0);
    __syncwarp();

    T* mps = reinterpret_cast<T*>(P.mps_out) + (size_t)img * P.mps_stride;
    int b = 1;
    int cols = cap;
    T* Pn = bufP;
    T* Qn = bufQ;
    T* Un = bufU;
    for (int n = 0; n < P.L; ++n) {
      const int q = cols >> 1;
      const int p = 2 * b;
      const bool rowmode = p <= q;
      const int NV = rowmode ? p : q;
      const int LEN = rowmode ? q : p;
      const int r = P.bonds[n + 1];
      T* V;
      if (rowmode) {
        V = Pn;  // vector j = 2l+s is the contiguous chunk j*q of Psi_n
        for (int i = lane; i < NV * NV; i += 32) Un[i] = (i / NV == i % NV) ? T(1) : T(0);
      } else {
        V = Qn;  // V[c][e = s*b + l] = Psi_n[l][s*q + c]
        for (int i = lane; i < NV * LEN; i += 32) {
          int c = i / LEN, e = i % LEN;
          int s = e / b, l = e % b;
          V[i] = Pn[l * cols + s * q + c];
        }
      }
      __syncwarp();
      // Frobenius norm -> zero threshold
      T fro = T(0);
      for (int i = lane; i < NV * LEN; i += 32) fro += V[i] * V[i];
      fro = warp_sum(fro);
      const T thr2 = Num<T>::floor_rel() * Num<T>::floor_rel() * fro;

      int sweeps = 0;
      if (NV > 1 && fro > T(0)) {
        for (; sweeps < 40; ++sweeps) {
          int rotated = 0;
          for (int i = 0; i < NV - 1; ++i) {
            for (int j = i + 1; j < NV; ++j) {
              T* a = V + i * LEN;
              T* bb_ = V + j * LEN;
              T aa = T(0), bb = T(0), ab = T(0);
              for (int e = lane; e < LEN; e += 32) {
                T x = a[e], y = bb_[e];
                aa += x * x;
                bb += y * y;
                ab += x * y;
              }
              aa = warp_sum(aa);
              bb = warp_sum(bb);
              ab = warp_sum(ab);
              if (aa > thr2 && bb > thr2 &&
                  Num<T>::abs(ab) > Num<T>::tol() * Num<T>::sqrt(aa) * Num<T>::sqrt(bb)) {
                T c, s;
                jacobi_cs(aa, bb, ab, c, s);
                for (int e = lane; e < LEN; e += 32) {
                  T x = a[e], y = bb_[e];
                  a[e] = c * x - s * y;
                  bb_[e] = s * x + c * y;
                }
                if (rowmode) {
                  T* ua = Un + i * NV;
                  T* ub = Un + j * NV;
                  for (int e = lane; e < NV; e += 32) {
                    T x = ua[e], y = ub[e];
                    ua[e] = c * x - s * y;
                    ub[e] = s * x + c * y;
                  }
                }
                rotated = 1;
                __syncwarp();
              }
            }
          }
          if (!rotated) break;
        }
      }
      // norms and stable descending rank; NV <= 32 -> one vector per lane
      T my = T(-1);
      if (lane < NV) {
        T acc = T(0);
        for (int e = 0; e < LEN; ++e) acc += V[lane * LEN + e] * V[lane * LEN + e];
        my = acc;
      }
      int rank = 0;
      for (int k = 0; k < NV; ++k) {
        T other = __shfl_sync(FULL, my, k);
        rank += (other > my) || (other == my && k < lane);
      }
      // order[rank] = lane
      int* order = reinterpret_cast<int*>(sig2 + 32);
      if (lane < NV) {
        order[rank] = lane;
        sig2[rank] = my;
      }
      __syncwarp();
      if (P.svals_out) {
        T* sv = reinterpret_cast<T*>(P.svals_out) + ((size_t)img * P.L + n) * P.sv_width;
        for (int k = lane; k < P.sv_width; k += 32)
          sv[k] = (k < NV && sig2[k] > thr2) ? Num<T>::sqrt(sig2[k]) : T(0);
      }
      if (P.sweeps_out && lane == 0) P.sweeps_out[(size_t)img * P.L + n] = sweeps;

      // A_n (2, b, r) row-major: element ((s*b + l)*r + k)
      T* A = mps + P.site_off[n];
      if (rowmode) {
        for (int i = lane; i < p * r; i += 32) {
          int k = i % r, e = i / r;  // e = s*b + l
          int s = e / b, l = e % b;
          T val = T(0);
          if (k < NV && sig2[k] > thr2) val = Un[order[k] * NV + 2 * l + s];
          A[i] = val;
        }
        // next Psi (r x q): rows = kept vectors in sorted order -> Qn
        for (int i = lane; i < r * q; i += 32) {
          int k = i / q, e = i % q;
          Qn[i] = (k < NV && sig2[k] > thr2) ? V[order[k] * LEN + e] : T(0);
        }
        __syncwarp();
        T* t = Pn; Pn = Qn; Qn = t;
      } else {
        // normalise kept columns in place (V = Qn), then Psi_next = U^T M
        for (int k = 0; k < r; ++k) {
          const bool keep = (k < NV) && sig2[k] > thr2;
          if (!keep) continue;
          T inv = Num<T>::rsqrt(sig2[k]);
          T* v = V + order[k] * LEN;
          for (int e = lane; e < LEN; e += 32) v[e] *= inv;
        }
        __syncwarp();
        for (int i = lane; i < p * r; i += 32) {
          int k = i % r, e = i / r;
          T val = T(0);
          if (k < NV && sig2[k] > thr2) val = V[order[k] * LEN + e];
          A[i] = val;
        }
        if (n + 1 < P.L) {
          for (int i = lane; i < r * q; i += 32) {
            int k = i / q, c = i % q;
            T acc = T(0);
            if (k < NV && sig2[k] > thr2) {
              const T* v = V + order[k] * LEN;
              for (int e = 0; e < LEN; ++e) {
                int s = e / b, l = e % b;
                acc += v[e] * Pn[l * cols + s * q + c];
              }
            }
            Un[i] = acc;
          }
          __syncwarp();
          T* t = Pn; Pn = Un; Un = t;
        }
      }
      __syncwarp();
      b = r;
      cols = q;
    }
    __syncwarp();
  }
}

// --------------------------------------------------------------------------
// Overlap: per image, left environment up to the hairy site c, right
// environment down to it, then the centre contraction with the label leg open
// (experiments.py:332-337).
// --------------------------------------------------------------------------
struct OverlapParams {
  const void* mps;
  const void* clf;
  int64_t N;
  int L;
  int c;  // hairy site
  int S;  // label-leg size
  int amax, Bmax;
  int64_t mps_stride;
  int64_t mps_off[OC_MAX_SITES + 1];
  int64_t clf_off[OC_MAX_SITES + 1];
  int a[OC_MAX_SITES + 1];
  int B[OC_MAX_SITES + 1];
  void* overlaps;
  int32_t* argmax;  // nullable
  int signed_out;   // 1: store the signed label vector instead of its absolute values (argmax still on |.|)
  // 1 (fast overlap kernel only): `mps` holds the chain of the BIT-REVERSED image, i.e. block position p is
  // (2, a_{L-p}, a_{L-1-p}) with element (s, r, l) = site L-1-p's (s, l, r) - the right-canonical MPS the
  // OC_TRUNC_SWEEP encoder has before its pass back to the left gauge (oc_sweep.cuh).  The overlaps do not
  // depend on the gauge, so the fused encode + classify entry points skip that pass.  Needs a[n] == a[L-n].
  int mirrored;
};

template <typename T>
__global__ void __launch_bounds__(128) overlap_generic_kernel(const __grid_constant__ OverlapParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int wpb = blockDim.x >> 5;
  const int eb = P.amax * P.Bmax;
  T* base = reinterpret_cast<T*>(smem_raw) + (size_t)warp * (5 * eb + 64);
  T* EL = base;            // a x B
  T* ER = base + eb;       // a x B
  T* X = base + 2 * eb;    // a x B ping-pong target
  T* Tm = base + 3 * eb;   // 2 x a x B scratch
  T* outv = base + 5 * eb; // S values
  const T* clf = reinterpret_cast<const T*>(P.clf);

  for (int64_t img = (int64_t)blockIdx.x * wpb + warp; img < P.N; img += (int64_t)gridDim.x * wpb) {
    const T* mps = reinterpret_cast<const T*>(P.mps) + (size_t)img * P.mps_stride;
    if (lane == 0) { EL[0] = T(1); ER[0] = T(1); }
    __syncwarp();
    // ---- left environment: sites 0..c-1 ----
    for (int n = 0; n < P.c; ++n) {
      const int al = P.a[n], ar = P.a[n + 1], Bl = P.B[n], Br = P.B[n + 1];
      const T* A = mps + P.mps_off[n];   // (2, al, ar)
      const T* W = clf + P.clf_off[n];   // (2, 1, Bl, Br)
      // Tm[s][a][B'] = sum_B EL[a][B] W[s][B][B']
      for (int i = lane; i < 2 * al * Br; i += 32) {
        int s = i / (al * Br), rem = i % (al * Br), a = rem / Br, Bp = rem % Br;
        T acc = T(0);
        for (int Bq = 0; Bq < Bl; ++Bq) acc += EL[a * Bl + Bq] * W[(s * Bl + Bq) * Br + Bp];
        Tm[i] = acc;
      }
      __syncwarp();
      // EL'[a'][B'] = sum_{s,a} A[s][a][a'] Tm[s][a][B']
      for (int i = lane; i < ar * Br; i += 32) {
        int ap = i / Br, Bp = i % Br;
        T acc = T(0);
        for (int s = 0; s < 2; ++s)
          for (int a = 0; a < al; ++a) acc += A[(s * al + a) * ar + ap] * Tm[(s * al + a) * Br + Bp];
        ER[i] = acc;  // temp: write into ER then swap
      }
      __syncwarp();
      T* t = EL; EL = ER; ER = t;
      __syncwarp();
    }
    if (lane == 0) ER[0] = T(1);
    __syncwarp();
    // ---- right environment: sites L-1..c+1 ----
    T* R = ER;
    T* R2 = X;
    for (int n = P.L - 1; n > P.c; --n) {
      const int al = P.a[n], ar = P.a[n + 1], Bl = P.B[n], Br = P.B[n + 1];
      const T* A = mps + P.mps_off[n];
      const T* W = clf + P.clf_off[n];
      // Tm[s][a][B'] = sum_a' A[s][a][a'] R[a'][B']
      for (int i = lane; i < 2 * al * Br; i += 32) {
        int s = i / (al * Br), rem = i % (al * Br), a = rem / Br, Bp = rem % Br;
        T acc = T(0);
        for (int ap = 0; ap < ar; ++ap) acc += A[(s * al + a) * ar + ap] * R[ap * Br + Bp];
        Tm[i] = acc;
      }
      __syncwarp();
      // R'[a][B] = sum_{s,B'} Tm[s][a][B'] W[s][B][B']
      for (int i = lane; i < al * Bl; i += 32) {
        int a = i / Bl, Bq = i % Bl;
        T acc = T(0);
        for (int s = 0; s < 2; ++s)
          for (int Bp = 0; Bp < Br; ++Bp) acc += Tm[(s * al + a) * Br + Bp] * W[(s * Bl + Bq) * Br + Bp];
        R2[i] = acc;
      }
      __syncwarp();
      T* t = R; R = R2; R2 = t;
    }
    // ---- centre site c ----
    {
      const int n = P.c;
      const int al = P.a[n], ar = P.a[n + 1], Bl = P.B[n], Br = P.B[n + 1];
      const T* A = mps + P.mps_off[n];   // (2, al, ar)
      const T* W = clf + P.clf_off[n];   // (2, S, Bl, Br)
      // F[s][B][a'] = sum_a EL[a][B] A[s][a][a']      -> Tm (2*Bl*ar <= 2*eb)
      for (int i = lane; i < 2 * Bl * ar; i += 32) {
        int s = i / (Bl * ar), rem = i % (Bl * ar), Bq = rem / ar, ap = rem % ar;
        T acc = T(0);
        for (int a = 0; a < al; ++a) acc += EL[a * Bl + Bq] * A[(s * al + a) * ar + ap];
        Tm[i] = acc;
      }
      __syncwarp();
      // H[s][B][B'] = sum_a' F[s][B][a'] R[a'][B'], folded straight into the
      // label sums out[l] = sum H[s][B][B'] W[s][l][B][B'] (H is never stored).
      T acc_l[OC_MAX_LABEL];
#pragma unroll
      for (int l = 0; l < OC_MAX_LABEL; ++l) acc_l[l] = T(0);
      for (int i = lane; i < 2 * Bl * Br; i += 32) {
        int s = i / (Bl * Br), rem = i % (Bl * Br), Bq = rem / Br, Bp = rem % Br;
        T h = T(0);
        for (int ap = 0; ap < ar; ++ap) h += Tm[(s * Bl + Bq) * ar + ap] * R[ap * Br + Bp];
#pragma unroll
        for (int l = 0; l < OC_MAX_LABEL; ++l)
          if (l < P.S) acc_l[l] += h * W[((size_t)(s * P.S + l) * Bl + Bq) * Br + Bp];
      }
#pragma unroll
      for (int l = 0; l < OC_MAX_LABEL; ++l) {
        if (l < P.S) {
          T v = warp_sum(acc_l[l]);
          if (lane == 0) outv[l] = v;
        }
      }
      __syncwarp();
      T* ov = reinterpret_cast<T*>(P.overlaps) + (size_t)img * P.S;
      for (int l = lane; l < P.S; l += 32) ov[l] = P.signed_out ? outv[l] : Num<T>::abs(outv[l]);
      if (P.argmax && lane == 0) {
        int best = 0;
        T bv = Num<T>::abs(outv[0]);
        for (int l = 1; l < P.S; ++l)
          if (Num<T>::abs(outv[l]) > bv) { bv = Num<T>::abs(outv[l]); best = l; }
        P.argmax[img] = best;
      }
    }
    __syncwarp();
  }
}

// out[i][m] (+)= | sum_j signed[i][j] * V[j][m] |  : the label legs of a one-hairy-site classifier
// contracted with M product bitstrings at once (variational_mpo_classifiers.py:313, 334-337:
// ``abs(img.H @ (clf @ bitstring))``; V[:, m] = bitstring m's vector on the hairy site times the
// product of its size-1 entries on the other sites).  n_sum > 1: the M columns are n_sum groups of
// n_out labels whose absolute values are added (the sum over paddings).
template <typename T>
__global__ void label_mix_kernel(const T* __restrict__ signed_ov, int64_t N, int S, const T* __restrict__ V,
                                 int n_out, int n_sum, T* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char mix_smem[];
  T* Vs = reinterpret_cast<T*>(mix_smem);
  const int M = n_out * n_sum;
  for (int i = threadIdx.x; i < S * M; i += blockDim.x) Vs[i] = V[i];
  __syncthreads();
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < N * n_out; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = e / n_out;
    const int l = (int)(e % n_out);
    const T* o = signed_ov + (size_t)i * S;
    T tot = T(0);
    for (int p = 0; p < n_sum; ++p) {
      T acc = T(0);
      for (int j = 0; j < S; ++j) acc += o[j] * Vs[j * M + p * n_out + l];
      tot += Num<T>::abs(acc);
    }
    out[e] = tot;
  }
}

// Ensemble voting on device (variational_mpo_classifiers.py:321-332, 348-363, k = 1):
// member predictions pred[k][i][0..n_labels) (|overlaps|, row stride S) -> every row normalised to
// sum 1 (``ensemble_predictions``, written to norm_out when given), their sum over members and its
// argmax (soft vote), and the label most members rank first (hard vote; ties go to the label whose
// first voter comes first, as collections.Counter.most_common does).
template <typename T>
__global__ void ensemble_vote_kernel(const T* __restrict__ pred, int K, int64_t N, int S, int n_labels,
                                     T* __restrict__ norm_out, T* __restrict__ soft_out,
                                     int32_t* __restrict__ soft_argmax, int32_t* __restrict__ hard_argmax) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    T soft[OC_MAX_LABEL];
    int votes[OC_MAX_LABEL], first[OC_MAX_LABEL];
#pragma unroll
    for (int l = 0; l < OC_MAX_LABEL; ++l) {
      soft[l] = T(0);
      votes[l] = 0;
      first[l] = 1 << 30;
    }
    for (int k = 0; k < K; ++k) {
      const T* p = pred + ((size_t)k * N + i) * S;
      T sum = T(0);
      for (int l = 0; l < n_labels; ++l) sum += p[l];
      const T inv = sum > T(0) ? T(1) / sum : T(0);
      int best = 0;
      T bv = p[0];
#pragma unroll
      for (int l = 0; l < OC_MAX_LABEL; ++l) {
        if (l < n_labels) {
          const T v = p[l] * inv;
          soft[l] += v;
          if (norm_out) norm_out[((size_t)k * N + i) * n_labels + l] = v;
          if (p[l] > bv) { bv = p[l]; best = l; }
        }
      }
#pragma unroll
      for (int l = 0; l < OC_MAX_LABEL; ++l) {
        if (l == best) {
          votes[l] += 1;
          first[l] = first[l] < k ? first[l] : k;
        }
      }
    }
    int sb = 0, hb = 0, hbv = -1, hbf = 1 << 30;
    T sbv = T(-1);
#pragma unroll
    for (int l = 0; l < OC_MAX_LABEL; ++l) {
      if (l < n_labels) {
        if (soft_out) soft_out[(size_t)i * n_labels + l] = soft[l];
        if (soft[l] > sbv) { sbv = soft[l]; sb = l; }
        if (votes[l] > hbv || (votes[l] == hbv && first[l] < hbf)) { hbv = votes[l]; hbf = first[l]; hb = l; }
      }
    }
    if (soft_argmax) soft_argmax[i] = sb;
    if (hard_argmax) hard_argmax[i] = hb;
  }
}

__global__ void count_correct_kernel(const int32_t* __restrict__ argmax,
                                     const uint8_t* __restrict__ labels, int64_t N,
                                     unsigned long long* __restrict__ out2) {
  unsigned long long local = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N;
       i += (int64_t)gridDim.x * blockDim.x)
    local += (argmax[i] == (int32_t)labels[i]);
  // warp reduce then one atomic per warp
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(FULL, local, o);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(out2, local);
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(out2 + 1, (unsigned long long)N);
}

}  // namespace oc
