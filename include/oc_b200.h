/*
 * oc_b200.h — C ABI of the B200-native encode + classify path of
 * Lewis1304/Orthogonal_Classifiers.
 *
 * The reference has no FFI: its hot path is plain Python.
 * Each entry point below names the reference code it replaces
 * (paths under /root/reference).  All `*_dev` pointers are CUDA device
 * pointers owned by the caller; every call is stream-ordered on `stream`
 * (a cudaStream_t passed as void*, NULL = default stream) and performs no
 * hidden synchronisation unless stated.  No torch types appear here.
 *
 * Library state.  Scratch memory, the re-laid classifier and auxiliary streams
 * live in a context keyed by (current device, stream): calls on different
 * streams or devices share nothing; host threads calling on the SAME stream are
 * serialised by a mutex.  Scratch of the device-pointer entry points grows with
 * cudaMallocAsync / cudaFreeAsync on the call's stream (stream-ordered, never a
 * device-wide synchronisation).  oc_release() frees a context.  The switches of
 * oc_set_option are process-global diagnostics, not per-call state.
 *
 * Return value: 0 = OK, non-zero = error (OC_E_*); the message is available
 * from oc_last_error() (thread-local).
 *
 * Array layouts (identical to the reference's):
 *   image        flat row-major 28x28 -> (n_pix,), tail zero-padded to 2^L
 *                inside the kernel (tools.py:218-220); site n <-> bit (L-1-n)
 *                of the flat index.
 *   MPS site n   (d=2, a_n, a_{n+1}) row-major, a_n = min(2^n, 2^(L-n), D)
 *                (fixed shape; rank-deficient trailing columns are zero, where
 *                xmps' minD would trim them).  Per image the L sites are
 *                concatenated: oc_mps_layout() gives bonds and offsets.
 *   classifier   site n (d=2, s_n, B_n, B_{n+1}) row-major (fMPO.py:21-51,
 *                tools.py:244-254), sites concatenated; exactly one site has
 *                s_n > 1 (the label leg, 16 for ten classes).
 */
#ifndef OC_B200_H
#define OC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OC_F32 0
#define OC_F64 1
#define OC_U8 2 /* input images only: value/255 is applied in-kernel */

#define OC_OK 0
#define OC_E_ARG 1      /* bad shape / unsupported size */
#define OC_E_IMAG 2     /* classifier has a non-zero imaginary part */
#define OC_E_CUDA 3     /* CUDA runtime error (message has the cudaError) */
#define OC_E_CAPACITY 4 /* output buffer too small */

/* encoder variants: which sweep applies the D truncation */
#define OC_TRUNC_INLINE 0 /* truncate inside the left-to-right SVD chain (BASELINE.json's wording) */
#define OC_TRUNC_SWEEP 1  /* exact chain, RIGHT-TO-LEFT truncating sweep, sweep back to the left-canonical
                           * gauge: the order of xmps' left_from_state(..).left_canonicalise(D) as called at
                           * tools.py:221 (restated in oracle/stubs/xmps/fMPS.py).  Identical to INLINE
                           * when D truncates no bond (D >= 2^floor(L/2), e.g. D >= 32 for 28x28 images). */

#define OC_MAX_L 10      /* n_pix <= 1024 */
#define OC_MAX_SITES 64  /* overlap kernel: chain length (MPS stacking) */
#define OC_MAX_LABEL 32  /* label-leg size */

int oc_version(void);
const char* oc_last_error(void);

/* Test/diagnostic switches (every alternative is CUDA; there is no CPU path).
 *   "force_generic" = 1  every call through the runtime-shaped kernels
 *   "enc_hw"   = 0  D >= 8 encodes with the one-image-per-warp kernels
 *   "enc_gram" = 0  steps 0-3 by Jacobi on the image unfoldings instead of the
 *                   16 x 16 Gram matrix
 *   "enc_pack" = 0  step 4 with 16 lanes per image instead of packed warps
 *   "enc_u8mma" = 0  uint8 pixels: G_3 = X_3 X_3^T through fp64 FMAs like float pixels (default 1: the exact
 *                   integer Gram matrix on the tensor cores, mma.sync u8 x u8 -> s32)
 *   "enc_mix"  = 0  packed step 4: only images with equally long lines share a warp (default 1: the
 *                   idle lanes of a warp also take one image with a shorter line; bit-identical results)
 *   "enc_mgs4" = 0  step 4 without the Gram-Schmidt preconditioning
 *   "enc_qr5"  = 0  step 5 without the pivoted Gram-Schmidt preconditioning
 *   "enc_qr6"  = 0  step 6 by Jacobi on the 32-element columns instead of Gram-Schmidt + the 8 x 8 factor
 *   "enc_split5" = 0  step 5 as one kernel instead of three
 *   "enc_sort" = 0  no device-side sort of the images by live rows
 *   "enc_split2" = 1  (opt-in) the two halves of an encoder chunk of >= 32,768 images
 *                   on two streams, rejoined on the caller's stream before return;
 *                   bit-identical results, ignored while "enc_events" is on
 *   "ov_static" = 0 run-time shapes in the overlap kernel for the 32/32 profile
 *   "ov_mma"   = 0  FFMA instead of tensor-core contractions in the overlap kernel
 *   "ov_pad"   = 0  D_encode < 32 images through the run-time-shaped overlap kernel
 *                   instead of zero-padded through the 32/32 one
 *   "ov_tc5"   = 1  (default 0) natural 32/32 profile: site 4 as one shared-operand tcgen05.mma
 *                   per CTA (images' A4^T E4 in tensor memory x classifier site in shared memory,
 *                   3xTF32, accumulator in TMEM) instead of per-image mma.sync; measured 5.8 %
 *                   slower than the default (profiles/tcgen05_r2.md), same overlaps
 *   "fused_sweep" = 0  oc_encode_classify / oc_classify_host with OC_TRUNC_SWEEP: bring the MPS back to
 *                   the left-canonical gauge before the overlap (as oc_encode_batch does) instead of
 *                   contracting the right-canonical MPS the right-to-left sweep leaves (same overlaps)
 *   "enc_events"    see oc_encode_step_ms
 *   "host_chunk" = n  images per pipeline stage of oc_classify_host (0 = default: 22,000 for uint8 and
 *                   8,750 for float pixels; with "host_streams" = 2: 17,500 / 8,750, = 1: 35,000 / 17,500)
 *   "host_streams" = 1 | 2 | 4  compute streams of oc_classify_host (default 4: chunk k runs on stream
 *                   k mod 4 - the caller's stream and three library-owned ones - so that the kernels of
 *                   several chunks fill each other's tail waves; 1: everything on the caller's stream;
 *                   results identical)
 *   "host_ramp" = 0 uint8 input: equal stages instead of the ramp (4 streams: 2/11, 4/11, 7/11 of a chunk, then
 *                   full chunks - 4,000 / 8,000 / 14,000 / 22,000 ...; 1 stream: chunk/4, chunk/2, chunk, ...)
 *   "host_trace" = 1  oc_classify_host prints a per-chunk timeline to stderr (2: also re-upload the classifier)
 * Unknown keys return OC_E_ARG. */
int oc_set_option(const char* key, int value);

/* Shape helper for the fixed-shape MPS batch.
 * bonds[L+1]: a_n;  site_offsets[L+1]: element offset of site n inside one
 * image's block, site_offsets[L] = elements per image (2728 for L=10, D=32).
 * D <= 0 means "no truncation" (the reference's D=None). */
int oc_mps_layout(int L, int D, int32_t* bonds, int64_t* site_offsets);

/* Width of one bond's row in svals_out: max_n min(2 a_n, 2^(L-n-1)). */
int oc_svals_width(int L, int D);

/* Replaces variational_mpo_classifiers.py:102-104 `mps_encoding` =
 * [fMPS_to_QTN(image_to_mps(img, D))] (tools.py:217-234): for each of N
 * images, pad to 2^L, run the successive truncated-SVD chain (xmps
 * left_from_state / left_canonicalise, call site tools.py:221) and write the
 * unit-norm left-canonical MPS.
 *   images_dev   N rows of n_pix values, row stride `ld` elements, in_dtype
 *                OC_F32 | OC_F64 | OC_U8
 *   dtype        arithmetic + output type, OC_F32 | OC_F64
 *   mps_out_dev  N * site_offsets[L] elements of `dtype`
 *   trunc_mode   OC_TRUNC_INLINE | OC_TRUNC_SWEEP
 *   svals_out_dev  nullable; N * L * oc_svals_width() elements: the singular
 *                values found at step n (bond n+1), descending, zero padded.
 *                OC_TRUNC_SWEEP: row n holds what the right-to-left sweep saw at
 *                bond n+1 (the image already truncated on the bonds to its right),
 *                row L-1 the norm of the truncated image
 *   sweeps_out_dev nullable; N * L int32 Jacobi sweep counts (diagnostics)
 * An all-zero image yields an all-zero MPS (the reference would fail). */
int oc_encode_batch(const void* images_dev, int in_dtype, int64_t N, int n_pix,
                    int64_t ld, int L, int D, int dtype, int trunc_mode,
                    void* mps_out_dev, void* svals_out_dev,
                    int32_t* sweeps_out_dev, void* stream);

/* Replaces the inline expression
 *   np.abs((mps_image.H.squeeze() @ classifier.squeeze()).data)
 * (experiments.py:278,332-337,390-391,439-440,633,769) and, restricted to
 * entries 0..n_labels-1, variational_mpo_classifiers.py:309-318
 * `classifier_predictions`; argmax_out is the top-1 of
 * `evaluate_classifier_top_k_accuracy` (:340-345).
 *   mps_dev       N image blocks laid out as oc_mps_layout (or any bonds:
 *                 mps_bonds_host[L+1], sites (2, a_n, a_{n+1}) concatenated)
 *   clf_dev       packed classifier (oc_pack_classifier layout) of `dtype`
 *   clf_bonds_host[L+1], clf_s_host[L]   host int arrays
 *   overlaps_out_dev  N * S elements (S = the label-leg size), |overlap|
 *   argmax_out_dev    nullable; N int32, first index of the maximum */
int oc_overlap_batch(const void* mps_dev, const int32_t* mps_bonds_host,
                     const void* clf_dev, const int32_t* clf_bonds_host,
                     const int32_t* clf_s_host, int L, int64_t N, int dtype,
                     void* overlaps_out_dev, int32_t* argmax_out_dev,
                     void* stream);

/* As oc_overlap_batch, but overlaps_out_dev receives the SIGNED label vector
 * (mps.H @ classifier).data before the np.abs (defined up to one global sign per
 * image, the sign of the MPS being a gauge choice); argmax_out is still the argmax
 * of the absolute values.  Input of oc_label_mix. */
int oc_overlap_batch_signed(const void* mps_dev, const int32_t* mps_bonds_host,
                            const void* clf_dev, const int32_t* clf_bonds_host,
                            const int32_t* clf_s_host, int L, int64_t N, int dtype,
                            void* overlaps_out_dev, int32_t* argmax_out_dev, void* stream);

/* Replaces the per-image, per-bitstring loops of
 * variational_mpo_classifiers.py:309-318 (classifier_predictions with arbitrary
 * product bitstrings) and :334-337 (padded_classifier_predictions):
 *   out[i][l] = sum_{p < n_sum} | sum_{j < S} signed[i][j] * V[j][p * n_out + l] |
 *   signed_dev  N x S signed label vectors (oc_overlap_batch_signed)
 *   V_dev       S x (n_out * n_sum), row-major: column m = bitstring m's vector on the
 *               hairy site times the product of its entries on the other sites
 *   out_dev     N x n_out */
int oc_label_mix(const void* signed_dev, int64_t N, int S, const void* V_dev, int n_out,
                 int n_sum, int dtype, void* out_dev, void* stream);

/* Replaces variational_mpo_classifiers.py:321-332 (ensemble_predictions: every
 * member's label vector normalised to sum 1) and the k = 1 case of :348-363
 * (evaluate_soft / evaluate_hard_ensemble_top_k_accuracy) up to the final count:
 *   pred_dev        K x N x S |overlaps| of the K members (labels 0..n_labels-1 used)
 *   norm_out_dev    nullable, K x N x n_labels normalised predictions
 *   soft_out_dev    nullable, N x n_labels sum over members of the normalised rows
 *   soft_argmax_dev nullable, N int32: argmax of that sum
 *   hard_argmax_dev nullable, N int32: the label most members rank first (ties: the
 *                   label whose first voter comes first, as Counter.most_common)
 * Feed the argmax arrays to oc_count_correct for the accuracies. */
int oc_ensemble_vote(const void* pred_dev, int K, int64_t N, int S, int n_labels, int dtype,
                     void* norm_out_dev, void* soft_out_dev, int32_t* soft_argmax_dev,
                     int32_t* hard_argmax_dev, void* stream);

/* encode + classify for one batch without the caller handling the MPS:
 * `workspace_dev` must hold N * site_offsets[L] elements of `dtype`
 * (pass NULL to let the library keep a grow-only internal workspace). */
int oc_encode_classify(const void* images_dev, int in_dtype, int64_t N,
                       int n_pix, int64_t ld, int L, int D, int dtype,
                       int trunc_mode, const void* clf_dev,
                       const int32_t* clf_bonds_host, const int32_t* clf_s_host,
                       void* workspace_dev, void* overlaps_out_dev,
                       int32_t* argmax_out_dev, void* stream);

/* ---- classifier construction ("batch adding"), fp64 on the device ----
 *
 * Replaces deterministic_mpo_classifier.py:21-42 (class_encode_mps_to_mpo / mpo_encoding)
 * for the reference's one-hot label bitstrings: MPO site n = MPS site n (label leg 1),
 * site c = A_c (x) e_label (label leg S).
 *   mps_dev        N image blocks, sites (2, a_n, a_{n+1}) concatenated (oc_encode_batch
 *                  output), OC_F32 or OC_F64
 *   labels_dev     N uint8 class labels
 *   mpos_out_dev   N blocks of sum_n 2 s_n a_n a_{n+1} doubles, s_c = S, s_n = 1 elsewhere */
int oc_class_encode(const void* mps_dev, int mps_dtype, int64_t N, int L,
                    const int32_t* mps_bonds_host, const uint8_t* labels_dev, int c, int S,
                    double* mpos_out_dev, void* stream);

/* Replaces fMPO.add + fMPO.compress_centre_one_site (+ scipy polar) — fMPO.py:666-689,
 * 338-453 — as adding_centre_batches / add_centre_sublist drive them
 * (experiments.py:497-523): groups of `batch` consecutive MPOs (the last may be shorter)
 * are summed and compressed to bond D around the hairy centre site c; `orthogonalise`
 * replaces the centre tensor by its polar factor, otherwise it is normalised.
 *   mpos_dev        n_mpos MPOs in one fixed zero-padded shape: sites (2, s_n, b_n, b_{n+1})
 *                   concatenated, bonds bonds_in_host[L+1] <= 32, s_c = S
 *   D               <= 0: the reference's D = None (numerical rank only)
 *   out_dev         ceil(n_mpos / batch) MPOs in the shape bonds_out_host describes;
 *                   NULL = only fill bonds_out_host (shape query)
 *   bonds_out_host  out: min(2 x previous, batch x input bond, D) from both edges */
int oc_add_compress_centre(const double* mpos_dev, int64_t n_mpos, int L, int c, int S,
                           const int32_t* bonds_in_host, int batch, int D, int orthogonalise,
                           double* out_dev, int32_t* bonds_out_host, void* stream);

/* Optional: declare the packed classifier at clf_dev immutable (enable = 1) so that
 * oc_overlap_batch / oc_encode_classify on this stream re-lay it for the kernel once
 * instead of on every call; enable = 1 again after changing its contents, enable = 0
 * before freeing it.  Without this call every overlap launch is preceded by a ~20 us
 * re-layout kernel. */
int oc_classifier_cache(const void* clf_dev, int enable, void* stream);

/* Free the scratch, cached classifiers and auxiliary streams of the context
 * (current device, stream).  Stream-ordered for the device-pointer scratch. */
int oc_release(void* stream);

/* The integer part of variational_mpo_classifiers.py:340-345 (k = 1):
 * out2_dev[0] += #(argmax == label), out2_dev[1] += N.  The caller zeroes
 * out2_dev; across ranks the two int64 are summed with one NCCL all-reduce. */
int oc_count_correct(const int32_t* argmax_dev, const uint8_t* labels_dev,
                     int64_t N, int64_t* out2_dev, void* stream);

/* Host-side, one-off: pack the reference's classifier container
 * (fMPO(data).data / [t.data for t in qtn.tensors], tools.py:244-254; the
 * per-site arrays load_qtn_classifier reads, tools.py:269-291) into one
 * contiguous buffer of `dtype`.
 *   site_ptrs_host[L]  host pointers to C-contiguous (2, s, Bl, Br) arrays of
 *                      float64 (is_complex = 0) or complex128 (is_complex = 1)
 *   shapes_host[4*L]   d, s, Bl, Br per site
 * Takes the real part after checking max|imag| == 0 (fMPO.add allocates
 * complex zeros, fMPO.py:668-679) -> OC_E_IMAG otherwise.
 *   packed_out_host    capacity elements of `dtype`; n_written = elements used
 *   bonds_out[L+1], s_out[L]  filled for oc_overlap_batch */
int oc_pack_classifier(const void* const* site_ptrs_host,
                       const int32_t* shapes_host, int L, int is_complex,
                       int dtype, void* packed_out_host, int64_t capacity,
                       int64_t* n_written, int32_t* bonds_out, int32_t* s_out);

/* End-to-end call with HOST buffers (what a reference user holds): images in
 * (float32 / float64 / uint8 as MNIST ships them, tools.py:25-74 divides by 255),
 * |overlaps| and argmax out.  The batch is cut into chunks; the host->device copy
 * of chunk k+1 (own copy stream) runs under the kernels of chunk k (`stream`) and
 * the results return on a third stream; the call returns after synchronising
 * `stream`.  Staging buffers are grow-only inside the context (plain cudaMalloc:
 * growing synchronises the device; sizes settle after the first call).  Copies
 * overlap only when the host buffers are pinned; pageable memory works, slower. */
int oc_classify_host(const void* images_host, int in_dtype, int64_t N,
                     int n_pix, int64_t ld, int L, int D, int dtype,
                     int trunc_mode, const void* clf_packed_host,
                     int64_t clf_elems, const int32_t* clf_bonds_host,
                     const int32_t* clf_s_host, void* overlaps_out_host,
                     int32_t* argmax_out_host, void* stream);

/* Diagnostics: per-launch device times of the encoder.  After
 * oc_set_option("enc_events", 1) every oc_encode_batch records a CUDA event on
 * its stream around the five parts of the step-major chain (Gram + steps 0-3 |
 * sort | step 4 | step 5 | steps 6-9; first chunk of the call only), keeping the
 * last 16 calls.  oc_encode_step_ms synchronises on those events and returns
 * the elapsed milliseconds of the `n` launches (n <= 5) averaged over the
 * recorded calls; used by bench.py for the roofline of the dominant kernel
 * (step 4).  Setting the option (to 0 or 1) forgets earlier records.  Returns
 * OC_E_ARG when nothing was recorded. */
int oc_encode_step_ms(float* ms_out_host, int n);

/* Diagnostics: peak-FMA microbenchmark kernel used by bench.py to measure the
 * fp32 FMA roofline denominator on the device it runs on.  Launches one
 * kernel of `iters` dependent-free FMA chains; returns flops executed.
 * iters < 0 runs |iters| iterations of the packed fma.rn.f32x2 variant. */
int oc_fma_peak_launch(int64_t iters, float* sink_dev, double* flops_out,
                       void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OC_B200_H */
