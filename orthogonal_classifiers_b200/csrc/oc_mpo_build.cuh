// Batch-added classifier construction on the device (fp64): fMPO.add + compress_centre_one_site
// (+ polar) of /root/reference/fMPO.py:338-453, 666-689 as driven by adding_centre_batches /
// add_centre_sublist (experiments.py:497-523), and the class encoding of
// deterministic_mpo_classifier.py:21-42.
//
// One CTA per GROUP of `batch` consecutive one-hairy-site MPOs.  The block-diagonal sum of
// fMPO.add is never materialised: with C the carry of the sweep (r x sum of the blocks' bonds),
//   left sweep, site m < c :  X_b[(d, rr), j] = sum_k C_b[rr][k] W_b[d][k][j]
//   right sweep, site m > c:  X_b[(d, rr), i] = sum_j W_b[d][i][j] C_b[rr][j]
// the matrix the reference hands to its SVD is M = [X_1 .. X_nb] (2r rows).  Its left singular
// vectors and singular values come from the SMALL Gram matrix G = sum_b X_b X_b^T (2r <= 64
// square, accumulated block by block in fp64; sigma = sqrt(lambda), so singular values below
// 1e-7 sigma_max are dropped where the reference keeps them down to 1e-14 - far below the 1e-5
// parity bar), eigen-decomposed by a CTA-wide one-sided Jacobi iteration in shared memory
// (round-robin ordering: p/2 disjoint pairs per round, one warp per pair).  New site = the
// eigenvectors of the D largest eigenvalues, new carry C' = U^T [X_1 .. X_nb], block by block.
// Centre: M_c[(d, rl), (s, rr)] = sum_b Cl_b W_b[d][s] Cr_b^T; normalised by its Frobenius norm
// (the sites either side are isometries, so that is sqrt(<O|O>), fMPO.py:446-450) or replaced by
// its polar factor (fMPO.py:438-444).  The polar factor weighs every singular direction equally,
// small ones included, so its SVD is done in two Gram passes (the second works on rows already
// orthogonal to ~1e-8 and graded, where Jacobi is accurate to the RELATIVE precision of each
// singular value); directions with sigma < 1e-10 sigma_max (the training images never excite
// them; scipy's completion there is arbitrary) are left out.
// Shapes are fixed per level: every MPO of a level has the same zero-padded bond profile.
#pragma once
#include "oc_generic.cuh"

namespace oc {
namespace build {

constexpr int kR = 32;        // largest bond in or out
constexpr int kP = 2 * kR;    // rows of a sweep matrix
constexpr int kLdG = kP + 1;  // leading dimension of G / V in shared memory
constexpr int kLdX = kR + 1;
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

struct BuildParams {
  const double* in;
  double* out;
  double* ws;
  int64_t n_mpos;
  int n_groups, batch;
  int L, c, S;
  int D;  // unused by the kernel (bo already holds min(D, caps)); kept for messages
  int ortho;
  int bi[OC_MAX_SITES + 1], bo[OC_MAX_SITES + 1];
  int64_t in_off[OC_MAX_SITES + 1], out_off[OC_MAX_SITES + 1];
  int64_t in_stride, out_stride, ws_stride;
  int Jmax;  // batch * max(bi)
};

// shared memory (doubles)
constexpr int kSmG = 0;
constexpr int kSmV = kSmG + kP * kLdG;
constexpr int kSmVt = kSmV + kP * kLdG;
constexpr int kSmX = kSmVt + kP * kLdG;
constexpr int kSmW = kSmX + kP * kLdX;
constexpr int kSmC = kSmW + 2 * kR * kLdX;
constexpr int kSmT = kSmC + kR * kLdX;
constexpr int kSmLam = kSmT + kR * kLdX;
constexpr int kSmTotal = kSmLam + kP + kP /* ord as ints fits */ + 8;
inline size_t smem_bytes() { return (size_t)kSmTotal * sizeof(double); }

// per-CTA workspace (doubles): 4 carries (left/right, ping-pong) of kR x Jmax, the centre matrix kP x (S kR)
inline int64_t ws_doubles(int Jmax, int S) { return 4LL * kR * Jmax + (int64_t)kP * S * kR; }

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// Eigen-decomposition of the symmetric positive semi-definite p x p matrix in Gs (column j at
// Gs + j * kLdG) by one-sided Jacobi on its columns: afterwards column j of Vs is an eigenvector,
// lam[j] its eigenvalue (the norm of column j of G V), ord[k] the index of the k-th largest.
__device__ void eig_sym(double* Gs, double* Vs, double* lam, int* ord, int p, int* flag) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int e = tid; e < p * kLdG; e += kThreads) {
    const int j = e / kLdG, r = e % kLdG;
    Vs[e] = (r == j) ? 1.0 : 0.0;
  }
  double fro = 0.0;
  for (int e = tid; e < p * p; e += kThreads) {
    const double v = Gs[(e / p) * kLdG + e % p];
    fro += v * v;
  }
  fro = warp_sum_d(fro);
  if (lane == 0) lam[warp] = fro;
  __syncthreads();
  fro = 0.0;
  for (int w = 0; w < kWarps; ++w) fro += lam[w];
  __syncthreads();
  const double thr2 = 1e-28 * fro;  // columns below 1e-14 |G|_F are numerically zero
  const int pe = p + (p & 1);       // a phantom column pairs with the odd one out
  if (p > 1) {
    for (int sweep = 0; sweep < 40; ++sweep) {
      if (tid == 0) *flag = 0;
      __syncthreads();
      for (int rd = 0; rd < pe - 1; ++rd) {
        for (int k = warp; k < pe / 2; k += kWarps) {
          int a, b;
          if (k == 0) {
            a = pe - 1;
            b = rd;
          } else {
            a = (rd + k) % (pe - 1);
            b = (rd - k + (pe - 1)) % (pe - 1);
          }
          if (a >= p || b >= p) continue;
          if (a > b) {
            const int t = a;
            a = b;
            b = t;
          }
          double* ga = Gs + a * kLdG;
          double* gb = Gs + b * kLdG;
          const double x0 = lane < p ? ga[lane] : 0.0, y0 = lane < p ? gb[lane] : 0.0;
          const double x1 = lane + 32 < p ? ga[lane + 32] : 0.0, y1 = lane + 32 < p ? gb[lane + 32] : 0.0;
          const double aa = warp_sum_d(x0 * x0 + x1 * x1), bb = warp_sum_d(y0 * y0 + y1 * y1),
                       ab = warp_sum_d(x0 * y0 + x1 * y1);
          if (aa > thr2 && bb > thr2 && ab * ab > 1e-30 * aa * bb) {
            double cs, sn;
            jacobi_cs(aa, bb, ab, cs, sn);
            if (lane < p) {
              ga[lane] = cs * x0 - sn * y0;
              gb[lane] = sn * x0 + cs * y0;
            }
            if (lane + 32 < p) {
              ga[lane + 32] = cs * x1 - sn * y1;
              gb[lane + 32] = sn * x1 + cs * y1;
            }
            double* va = Vs + a * kLdG;
            double* vb = Vs + b * kLdG;
            if (lane < p) {
              const double u = va[lane], w = vb[lane];
              va[lane] = cs * u - sn * w;
              vb[lane] = sn * u + cs * w;
            }
            if (lane + 32 < p) {
              const double u = va[lane + 32], w = vb[lane + 32];
              va[lane + 32] = cs * u - sn * w;
              vb[lane + 32] = sn * u + cs * w;
            }
            if (lane == 0) *flag = 1;
          }
        }
        __syncthreads();
      }
      const int again = *flag;
      __syncthreads();
      if (!again) break;
    }
  }
  for (int j = warp; j < p; j += kWarps) {
    const double* g = Gs + j * kLdG;
    const double x0 = lane < p ? g[lane] : 0.0, x1 = lane + 32 < p ? g[lane + 32] : 0.0;
    const double n2 = warp_sum_d(x0 * x0 + x1 * x1);
    if (lane == 0) lam[j] = sqrt(n2);
  }
  __syncthreads();
  if (tid < p) {
    const double mine = lam[tid];
    int rank = 0;
    for (int j = 0; j < p; ++j) {
      const double o = lam[j];
      rank += (o > mine) || (o == mine && j < tid);
    }
    ord[rank] = tid;
  }
  __syncthreads();
}

// X_b of one block into Xs (p = 2 r rows, ncol columns); W_b and C_b staged in Ws / Cs
template <bool LEFT>
__device__ void build_x(const BuildParams& P, int m, int64_t mpo, const double* __restrict__ carry, int Jc, int b, int r,
                        double* Xs, double* Ws, double* Cs) {
  const int tid = threadIdx.x;
  const int ib = P.bi[m], jb = P.bi[m + 1];
  const double* W = P.in + mpo * P.in_stride + P.in_off[m];  // (2, 1, ib, jb)
  for (int e = tid; e < 2 * ib * jb; e += kThreads) {
    const int jj = e % jb, dk = e / jb, k = dk % ib, d = dk / ib;
    Ws[(d * kR + k) * kLdX + jj] = W[e];
  }
  const int nc = LEFT ? ib : jb;  // carry columns of this block
  for (int e = tid; e < r * nc; e += kThreads) {
    const int k = e % nc, rr = e / nc;
    Cs[rr * kLdX + k] = carry[(size_t)rr * Jc + b * nc + k];
  }
  __syncthreads();
  const int ncol = LEFT ? jb : ib;
  for (int e = tid; e < 2 * r * ncol; e += kThreads) {
    const int col = e % ncol, row = e / ncol, d = row / r, rr = row % r;
    double acc = 0.0;
    if (LEFT) {
      for (int k = 0; k < ib; ++k) acc += Cs[rr * kLdX + k] * Ws[(d * kR + k) * kLdX + col];
    } else {
      for (int j = 0; j < jb; ++j) acc += Ws[(d * kR + col) * kLdX + j] * Cs[rr * kLdX + j];
    }
    Xs[row * kLdX + col] = acc;
  }
  __syncthreads();
}

// One site of a sweep: Gram matrix over the blocks, eigenvectors -> out site, new carry.
template <bool LEFT>
__device__ void sweep_site(const BuildParams& P, int m, int64_t first, int nb, const double* carry_cur, double* carry_next,
                           double* out_mpo, double* sm, int* flag) {
  const int tid = threadIdx.x;
  double* Gs = sm + kSmG;
  double* Vs = sm + kSmV;
  double* Xs = sm + kSmX;
  double* Ws = sm + kSmW;
  double* Cs = sm + kSmC;
  double* lam = sm + kSmLam;
  int* ord = reinterpret_cast<int*>(lam + kP);
  const int r = LEFT ? P.bo[m] : P.bo[m + 1];   // rank carried into this site
  const int rn = LEFT ? P.bo[m + 1] : P.bo[m];  // rank kept
  const int p = 2 * r;
  const int ncol = LEFT ? P.bi[m + 1] : P.bi[m];      // columns of X_b = the blocks' far bond
  const int Jc = nb * (LEFT ? P.bi[m] : P.bi[m + 1]);  // columns of the current carry
  const int Jn = nb * ncol;
  // ---- G = sum_b X_b X_b^T: thread (ti, tj) owns a 4 x 4 tile ----
  const int ti = tid >> 4, tj = tid & 15;
  double acc[4][4];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) acc[u][v] = 0.0;
  for (int b = 0; b < nb; ++b) {
    build_x<LEFT>(P, m, first + b, carry_cur, Jc, b, r, Xs, Ws, Cs);
    if (4 * ti < p && 4 * tj < p) {
      for (int col = 0; col < ncol; ++col) {
        double xa[4], xb[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          xa[u] = 4 * ti + u < p ? Xs[(4 * ti + u) * kLdX + col] : 0.0;
          xb[u] = 4 * tj + u < p ? Xs[(4 * tj + u) * kLdX + col] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int v = 0; v < 4; ++v) acc[u][v] += xa[u] * xb[v];
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v)
      if (4 * ti + u < p && 4 * tj + v < p) Gs[(4 * tj + v) * kLdG + 4 * ti + u] = acc[u][v];
  __syncthreads();
  eig_sym(Gs, Vs, lam, ord, p, flag);
  const double lmax = lam[ord[0]];
  const double lthr = 1e-14 * lmax;
  // ---- site: LEFT (2, 1, r, rn): A[d][rr][k] = V[(d, rr)][k];  RIGHT (2, 1, rn, r): A[d][k][rr] ----
  {
    double* A = out_mpo + P.out_off[m];
    for (int e = tid; e < 2 * r * rn; e += kThreads) {
      int d, rr, k;
      if (LEFT) {
        k = e % rn;
        rr = (e / rn) % r;
        d = e / (rn * r);
      } else {
        rr = e % r;
        k = (e / r) % rn;
        d = e / (r * rn);
      }
      const bool keep = k < p && lam[ord[k < p ? k : 0]] > lthr && lmax > 0.0;
      A[e] = keep ? Vs[ord[k] * kLdG + d * r + rr] : 0.0;
    }
  }
  // ---- new carry C'[k][b * ncol + col] = sum_row V[row][k] X_b[row][col] ----
  for (int b = 0; b < nb; ++b) {
    build_x<LEFT>(P, m, first + b, carry_cur, Jc, b, r, Xs, Ws, Cs);
    for (int e = tid; e < rn * ncol; e += kThreads) {
      const int col = e % ncol, k = e / ncol;
      double a = 0.0;
      if (k < p && lam[ord[k]] > lthr && lmax > 0.0) {
        const double* v = Vs + ord[k] * kLdG;
        for (int row = 0; row < p; ++row) a += v[row] * Xs[row * kLdX + col];
      }
      carry_next[(size_t)k * Jn + b * ncol + col] = a;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kThreads, 1) add_compress_centre_kernel(const __grid_constant__ BuildParams P) {
  extern __shared__ __align__(16) double sm[];
  __shared__ int flag;
  const int tid = threadIdx.x;
  double* ws = P.ws + (size_t)blockIdx.x * P.ws_stride;
  double* cl[2] = {ws, ws + (size_t)kR * P.Jmax};
  double* cr[2] = {ws + 2 * (size_t)kR * P.Jmax, ws + 3 * (size_t)kR * P.Jmax};
  double* Mc = ws + 4 * (size_t)kR * P.Jmax;
  double* Gs = sm + kSmG;
  double* Vs = sm + kSmV;
  double* Vt = sm + kSmVt;
  double* Xs = sm + kSmX;
  double* Ws = sm + kSmW;
  double* Cs = sm + kSmC;
  double* Ts = sm + kSmT;
  double* lam = sm + kSmLam;
  int* ord = reinterpret_cast<int*>(lam + kP);
  const int c = P.c, L = P.L, S = P.S;
  for (int g = blockIdx.x; g < P.n_groups; g += gridDim.x) {
    const int64_t first = (int64_t)g * P.batch;
    const int nb = (int)((P.n_mpos - first) < P.batch ? (P.n_mpos - first) : P.batch);
    double* out_mpo = P.out + (size_t)g * P.out_stride;
    // ---- left sweep: the first site concatenates the blocks (carry of ones), fMPO.py:668-671 ----
    for (int e = tid; e < nb; e += kThreads) {
      cl[0][e] = 1.0;
      cr[0][e] = 1.0;
    }
    __syncthreads();
    int pl = 0;
    for (int m = 0; m < c; ++m) {
      sweep_site<true>(P, m, first, nb, cl[pl], cl[pl ^ 1], out_mpo, sm, &flag);
      pl ^= 1;
    }
    int pr = 0;
    for (int m = L - 1; m > c; --m) {
      sweep_site<false>(P, m, first, nb, cr[pr], cr[pr ^ 1], out_mpo, sm, &flag);
      pr ^= 1;
    }
    // ---- centre matrix M_c[(d, rl)][(s, rr)] = sum_b Cl_b W_b[d][s] Cr_b^T ----
    const int rl = P.bo[c], rr_n = P.bo[c + 1], ib = P.bi[c], jb = P.bi[c + 1];
    const int ncols = S * rr_n, p = 2 * rl;
    for (int e = tid; e < p * ncols; e += kThreads) Mc[e] = 0.0;
    __syncthreads();
    for (int b = 0; b < nb; ++b) {
      // Cl_b -> Cs (rl x ib), Cr_b -> Xs (rr_n x jb)
      for (int e = tid; e < rl * ib; e += kThreads) Cs[(e / ib) * kLdX + e % ib] = cl[pl][(size_t)(e / ib) * (nb * ib) + b * ib + e % ib];
      for (int e = tid; e < rr_n * jb; e += kThreads) Xs[(e / jb) * kLdX + e % jb] = cr[pr][(size_t)(e / jb) * (nb * jb) + b * jb + e % jb];
      const double* Wc = P.in + (first + b) * P.in_stride + P.in_off[c];  // (2, S, ib, jb)
      for (int ds = 0; ds < 2 * S; ++ds) {
        __syncthreads();
        int nz = 0;
        for (int e = tid; e < ib * jb; e += kThreads) {
          const double v = Wc[(size_t)ds * ib * jb + e];
          Ws[(e / jb) * kLdX + e % jb] = v;
          nz |= v != 0.0;
        }
        if (!__syncthreads_or(nz)) continue;  // class-encoded MPOs carry one label slice
        // T[rl][j] = sum_k Cl[rl][k] W[k][j]
        for (int e = tid; e < rl * jb; e += kThreads) {
          const int j = e % jb, a = e / jb;
          double acc = 0.0;
          for (int k = 0; k < ib; ++k) acc += Cs[a * kLdX + k] * Ws[k * kLdX + j];
          Ts[a * kLdX + j] = acc;
        }
        __syncthreads();
        const int d = ds / S, s = ds % S;
        for (int e = tid; e < rl * rr_n; e += kThreads) {
          const int q = e % rr_n, a = e / rr_n;
          double acc = 0.0;
          for (int j = 0; j < jb; ++j) acc += Ts[a * kLdX + j] * Xs[q * kLdX + j];
          Mc[(size_t)(d * rl + a) * ncols + s * rr_n + q] += acc;
        }
      }
      __syncthreads();
    }
    double* Ac = out_mpo + P.out_off[c];  // (2, S, rl, rr_n)
    if (!P.ortho) {
      double n2 = 0.0;
      for (int e = tid; e < p * ncols; e += kThreads) n2 += Mc[e] * Mc[e];
      n2 = warp_sum_d(n2);
      if ((tid & 31) == 0) lam[tid >> 5] = n2;
      __syncthreads();
      n2 = 0.0;
      for (int w = 0; w < kWarps; ++w) n2 += lam[w];
      __syncthreads();
      const double inv = n2 > 0.0 ? 1.0 / sqrt(n2) : 0.0;
      for (int e = tid; e < p * ncols; e += kThreads) {
        const int col = e % ncols, row = e / ncols, d = row / rl, a = row % rl, s = col / rr_n, q = col % rr_n;
        Ac[((size_t)(d * S + s) * rl + a) * rr_n + q] = Mc[e] * inv;
      }
    } else {
      // ---- polar factor of M_c (p x ncols, p <= ncols): two Gram passes, Y <- V^T Y in place ----
      for (int e = tid; e < p * kLdG; e += kThreads) Vt[e] = (e % kLdG == e / kLdG) ? 1.0 : 0.0;  // V_total, column-major
      __syncthreads();
      const int ti = tid >> 4, tj = tid & 15;
      for (int pass = 0; pass < 2; ++pass) {
        double acc[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int v = 0; v < 4; ++v) acc[u][v] = 0.0;
        for (int c0 = 0; c0 < ncols; c0 += kR) {
          const int w = ncols - c0 < kR ? ncols - c0 : kR;
          for (int e = tid; e < p * w; e += kThreads) Xs[(e / w) * kLdX + e % w] = Mc[(size_t)(e / w) * ncols + c0 + e % w];
          __syncthreads();
          if (4 * ti < p && 4 * tj < p) {
            for (int col = 0; col < w; ++col) {
              double xa[4], xb[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                xa[u] = 4 * ti + u < p ? Xs[(4 * ti + u) * kLdX + col] : 0.0;
                xb[u] = 4 * tj + u < p ? Xs[(4 * tj + u) * kLdX + col] : 0.0;
              }
#pragma unroll
              for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < 4; ++v) acc[u][v] += xa[u] * xb[v];
            }
          }
          __syncthreads();
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int v = 0; v < 4; ++v)
            if (4 * ti + u < p && 4 * tj + v < p) Gs[(4 * tj + v) * kLdG + 4 * ti + u] = acc[u][v];
        __syncthreads();
        eig_sym(Gs, Vs, lam, ord, p, &flag);
        // Y <- V^T Y (rows in eigenvalue order), chunk by chunk
        for (int c0 = 0; c0 < ncols; c0 += kR) {
          const int w = ncols - c0 < kR ? ncols - c0 : kR;
          for (int e = tid; e < p * w; e += kThreads) Xs[(e / w) * kLdX + e % w] = Mc[(size_t)(e / w) * ncols + c0 + e % w];
          __syncthreads();
          for (int e = tid; e < p * w; e += kThreads) {
            const int col = e % w, k = e / w;
            const double* v = Vs + ord[k] * kLdG;
            double a = 0.0;
            for (int row = 0; row < p; ++row) a += v[row] * Xs[row * kLdX + col];
            Mc[(size_t)k * ncols + c0 + col] = a;
          }
          __syncthreads();
        }
        // V_total <- V_total V[:, ord]  (through Gs as scratch)
        for (int e = tid; e < p * p; e += kThreads) {
          const int row = e % p, k = e / p;
          const double* v = Vs + ord[k] * kLdG;
          double a = 0.0;
          for (int q = 0; q < p; ++q) a += Vt[q * kLdG + row] * v[q];
          Gs[k * kLdG + row] = a;
        }
        __syncthreads();
        for (int e = tid; e < p * p; e += kThreads) Vt[(e / p) * kLdG + e % p] = Gs[(e / p) * kLdG + e % p];
        __syncthreads();
      }
      // rows of Y are orthogonal with norms sigma_k = sqrt(lambda_k) (descending):
      // U_polar = V_total diag(1 / sigma) Y over the directions with sigma > 1e-10 sigma_max
      const double smax = sqrt(lam[ord[0]]);
      if (tid < p) {
        const double sg = sqrt(lam[ord[tid]]);
        Ts[tid] = (smax > 0.0 && sg > 1e-10 * smax) ? 1.0 / sg : 0.0;
      }
      __syncthreads();
      for (int c0 = 0; c0 < ncols; c0 += kR) {
        const int w = ncols - c0 < kR ? ncols - c0 : kR;
        for (int e = tid; e < p * w; e += kThreads)
          Xs[(e / w) * kLdX + e % w] = Mc[(size_t)(e / w) * ncols + c0 + e % w] * Ts[e / w];
        __syncthreads();
        for (int e = tid; e < p * w; e += kThreads) {
          const int cc = e % w, row = e / w;
          double a = 0.0;
          for (int k = 0; k < p; ++k) a += Vt[k * kLdG + row] * Xs[k * kLdX + cc];
          const int col = c0 + cc, d = row / rl, aa = row % rl, s = col / rr_n, q = col % rr_n;
          Ac[((size_t)(d * S + s) * rl + aa) * rr_n + q] = a;
        }
        __syncthreads();
      }
    }
    __syncthreads();
  }
}

// deterministic_mpo_classifier.py:21-42 (class_encode_mps_to_mpo / mpo_encoding) for one-hot label
// bitstrings on the hairy site c: MPO site n = MPS site n (label leg 1), site c = A_c (x) e_label.
template <typename TIN>
__global__ void class_encode_kernel(const TIN* __restrict__ mps, int64_t mps_stride, const uint8_t* __restrict__ labels,
                                    int64_t N, int S, int c_lo, int c_hi /* element range of site c inside an MPS */,
                                    double* __restrict__ out, int64_t out_stride) {
  const int64_t csz = c_hi - c_lo;
  for (int64_t img = blockIdx.x; img < N; img += gridDim.x) {
    const TIN* a = mps + (size_t)img * mps_stride;
    double* o = out + (size_t)img * out_stride;
    const int label = labels[img];
    // sites before c | site c expanded (2, S, a, a') | sites after c
    for (int64_t e = threadIdx.x; e < c_lo; e += blockDim.x) o[e] = (double)a[e];
    const int64_t half = csz / 2;  // elements of one physical slice of site c
    for (int64_t e = threadIdx.x; e < csz * S; e += blockDim.x) {
      const int64_t within = e % half, ds = e / half, s = ds % S, d = ds / S;
      o[c_lo + e] = s == label ? (double)a[c_lo + d * half + within] : 0.0;
    }
    for (int64_t e = c_hi + threadIdx.x; e < mps_stride; e += blockDim.x) o[e + csz * (S - 1)] = (double)a[e];
  }
}

}  // namespace build
}  // namespace oc
