/*
 * vidseg_b200.h -- C-ABI of libvidseg_b200.so (sm_100a).
 *
 * Drop-in boundary for the per-clip hot path of QianWangX/VidSeg_diffusion.
 * The reference is pure Python on PyTorch/scikit-learn and has no FFI of its
 * own; each entry point below names the reference call site (file:line under
 * the reference root) whose arithmetic it replaces.  INTEGRATION.md shows the
 * ctypes stubs a maintainer adds on the reference side.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - tensors are contiguous, row-major, dtype as declared;
 *   - the caller owns all memory (outputs and workspaces are caller-allocated);
 *   - `stream` is a cudaStream_t passed as void*; kernels are enqueued on it and
 *     the call returns without synchronising unless stated otherwise;
 *   - return value 0 = success, otherwise a cudaError_t or a negative VIDSEG_E_*
 *     code; vidseg_last_error() gives the message for the calling thread.
 */
#ifndef VIDSEG_B200_H_
#define VIDSEG_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VIDSEG_E_INVALID   (-1)  /* bad argument (shape, alignment, null pointer) */
#define VIDSEG_E_WORKSPACE (-2)  /* workspace too small                         */
#define VIDSEG_E_UNSUPPORTED (-3)

const char* vidseg_last_error(void);
/* ABI version of this header; bumped on any signature change. */
int vidseg_abi_version(void);  /* 7: + first-stage entries (conv2d_down_pad_after, softmax_rows), seg-map post-process; 6: + kmeans exchange words, flags_async, aggregate_normalize_rows; 5: + sampler_step, set/get_kmeans_mstep; 4: + row_scalar in gemm_split_ex, gemm_geglu_split, majority_map / knn_predict (3: operand policy, split_rows) */
/* Compute capability major*10+minor of the current device (100 on B200). */
int vidseg_device_arch(void);

/* ------------------------------------------------------------------------- *
 * R1  feature aggregation + per-token max-abs normalisation
 *
 * Replaces scripts/sampling/feature_extraction.py:739-745 (torch.mean over the
 * stacked blocks, in caller order), :38-39 (x / max|x| over channels, no
 * epsilon) and :45-46 (keep the conditional half rows [F, 2F), flatten).
 *   blocks[b] : float32 [2F, hw, C], b < n_blocks (1..4)
 *   out       : float32 [F*hw, C]
 * ------------------------------------------------------------------------- */
int vidseg_aggregate_normalize(const float* const* blocks_host, int n_blocks,
                               int num_frames, int hw, int channels,
                               float* out, void* stream);

/* Same arithmetic on the row window [first_row, first_row + rows) of every block (blocks are [*, C] row-major): the
 * CFG-half split of the SVD UNet over two GPUs leaves the conditional rows alone on one rank (first_row = 0). */
int vidseg_aggregate_normalize_rows(const float* const* blocks_host, int n_blocks,
                                    long long first_row, long long rows, int channels,
                                    float* out, void* stream);

/* ------------------------------------------------------------------------- *
 * R2  K-means  (sklearn.cluster.KMeans(n_clusters=K, n_init=R).fit + .predict)
 *
 * Replaces scripts/sampling/feature_extraction.py:52-55.  The random decisions
 * of k-means++ are data-independent draws from numpy's global RandomState; the
 * host mirror draws them (same calls, same order as sklearn/_kmeans.py:231,249)
 * and passes them in:
 *   first_idx : int32  [R]            first centre of every run (choice(n, p))
 *   rand      : double [R, K-1, T]    uniform(size=T) draws, T = 2 + int(ln K)
 * All R runs advance in lock-step on the device (batched seeding, batched Lloyd).
 * ------------------------------------------------------------------------- */
size_t vidseg_kmeans_workspace_bytes(int n, int d, int k, int n_init, int n_trials);

/* Variant of the Lloyd M-step (sklearn/cluster/_k_means_lloyd.pyx:_update_chunk_dense, the centers_new sums).  Process
 * wide; read when a workspace is sized and prepared, so set it before vidseg_kmeans_workspace_bytes.
 *   1  (default; VIDSEG_KMEANS_MSTEP=0 in the environment selects 0) the sums of all runs are one int8 tensor-core
 *      product: rows carried as 48-bit fixed point in six signed base-256 digit planes, labels as a 0/1 matrix,
 *      tcgen05.mma.kind::i8 accumulating in int32 -- exact integer arithmetic, independent of tiling and order;
 *   0  float64 accumulation in shared memory in a fixed slab order.
 * Both produce the same `partial` [R, K, D+1] float64 array up to the last bit of float64 rounding. */
int vidseg_set_kmeans_mstep(int mode);
int vidseg_get_kmeans_mstep(void);

/* mean-centre (sequential fp32 column sums, as numpy's X.mean(axis=0)),
 * tolerance mean(var(X,0))*tol_rel, float64 squared row norms; registers the
 * problem shape with the workspace (host side) and resets the per-run state. */
int vidseg_kmeans_prepare(const float* x, int n, int d, int k, int n_init, int n_trials,
                          float tol_rel, int max_iter, void* workspace, size_t workspace_bytes,
                          void* stream);

/* k-means++ seeding of all runs (sklearn/_kmeans.py:180-277). */
int vidseg_kmeans_seed(const int32_t* first_idx, const double* rand,
                       void* workspace, size_t workspace_bytes, void* stream);

/* E-step over rows [row_begin,row_end): labels = argmin_j ||c_j||^2 - 2 x.c_j
 * (sklearn/_k_means_lloyd.pyx:_update_chunk_dense) for every unfinished run. */
int vidseg_kmeans_assign(void* workspace, size_t workspace_bytes, int row_begin, int row_end,
                         void* stream);

/* M-step part 1: per-run, per-cluster float64 sums and counts of rows
 * [row_begin,row_end) -> partial [R, K, D+1] (column D = count) and changed
 * int32 [R] (labels that changed in the last assign).  NULL = keep them in the
 * workspace.  In the multi-GPU path the caller all-reduces both between part 1
 * and part 2. */
int vidseg_kmeans_partial(void* workspace, size_t workspace_bytes, int row_begin, int row_end,
                          double* partial, int32_t* changed, void* stream);

/* M-step part 2: relocate empty clusters (modifies `partial` in place; skipped
 * when local_rows_only != 0 because it needs every label), average, centre
 * shift, convergence flags (strict label equality, else sum(shift^2) <= tol). */
int vidseg_kmeans_update(void* workspace, size_t workspace_bytes, double* partial,
                         const int32_t* changed, int local_rows_only, void* stream);

/* Multi-GPU exchange form of one Lloyd iteration ("allreduce inside K-means"): the per-cluster
 * sums, the counts and the label-change counters of all runs travel as ONE array of 8-byte words
 *   words[(r*K + j)*(D+1) + c] sums (c < D) | count (c == D),  then words[R*K*(D+1) + r] = changed labels of run r,
 * so an iteration costs one all-reduce.  words_i64 != 0: the words are the INTEGER fixed-point sums of the int8
 * tensor-core M-step (int64) -- all-reduce(SUM) over ranks is then exact and order independent and the sharded fit is
 * bit-identical to the single-GPU fit; words_i64 == 0: float64 words.  vidseg_kmeans_exchange_mode returns 1 when the
 * row range of a rank can produce integer words (every rank must pass the same words_i64, so the caller ANDs the
 * answers for all ranks' ranges), 0 when not, negative on error.  update_words with integer words requires
 * local_rows_only (no empty-cluster relocation in that form; see vidseg_kmeans_status). */
size_t vidseg_kmeans_exchange_words(void* workspace, size_t workspace_bytes);
int vidseg_kmeans_exchange_mode(void* workspace, size_t workspace_bytes, int row_begin, int row_end);
int vidseg_kmeans_partial_words(void* workspace, size_t workspace_bytes, int row_begin, int row_end,
                                int words_i64, void* words, void* stream);
int vidseg_kmeans_update_words(void* workspace, size_t workspace_bytes, int words_i64, void* words,
                               int local_rows_only, void* stream);
/* convergence flags int32 [R][4] = {done, strict, n_iter, empty cluster seen} -> (pinned) host memory WITHOUT
 * synchronising: record an event behind the call and read the flags once it has passed. */
int vidseg_kmeans_flags_async(void* workspace, size_t workspace_bytes, int32_t* flags_host, void* stream);

/* number of runs still iterating (synchronises the stream). */
int vidseg_kmeans_active_runs(void* workspace, size_t workspace_bytes, int* n_active_host,
                              void* stream);

/* as active_runs, plus the number of runs whose M-step met an EMPTY cluster while local_rows_only was set (the
 * relocation of sklearn/_k_means_common.pyx:_relocate_empty_clusters_dense needs every label, so it is skipped
 * there): a sharded caller that sees empty_seen != 0 repeats the fit unsharded.  Synchronises the stream. */
int vidseg_kmeans_status(void* workspace, size_t workspace_bytes, int* n_active_host, int* empty_seen_host,
                         void* stream);

/* final E-step of the runs that did not converge strictly, then the inertia of
 * rows [row_begin,row_end) -> inertia_partial double [R] (NULL = workspace). */
int vidseg_kmeans_inertia(void* workspace, size_t workspace_bytes, int row_begin, int row_end,
                          double* inertia_partial, void* stream);

/* same[a*R+b] = 1 iff labels of run a map onto labels of run b as a function on
 * rows [row_begin,row_end) (_k_means_common.pyx:_is_same_clustering). */
int vidseg_kmeans_same_matrix(void* workspace, size_t workspace_bytes, int row_begin, int row_end,
                              int32_t* same, void* stream);

/* best-of-R rule on the host (sklearn/_kmeans.py:1529-1541). */
int vidseg_kmeans_pick_best_host(const float* inertia_host, const int32_t* same_host, int n_init);

/* un-centre run `best` -> centers_out float32 [K, D]; labels_fit_out int32 [N] or NULL. */
int vidseg_kmeans_finish(void* workspace, size_t workspace_bytes, int best, float* centers_out,
                         int32_t* labels_fit_out, void* stream);

/* single-GPU convenience: same_matrix + pick_best + finish.  info_host int32[4] =
 * {best run, n_iter of best, max n_iter, 0}; inertia_host float[R].  Synchronises. */
int vidseg_kmeans_select(void* workspace, size_t workspace_bytes, const double* inertia,
                         float* centers_out, int32_t* labels_fit_out,
                         int32_t* info_host, float* inertia_host, void* stream);

/* KMeans.predict: argmin over un-centred x against centres float32 [K, D].
 * cnorm_scratch: double [K]. */
int vidseg_kmeans_predict(const float* x, int n, int d, const float* centers, int k,
                          int32_t* labels_out, double* cnorm_scratch, void* stream);

/* `iterations` Lloyd iterations of every unfinished run over all rows, queued on `stream` without synchronising: the
 * loop body of vidseg_kmeans_fit_predict (sklearn/cluster/_kmeans.py:_kmeans_single_lloyd, one call of lloyd_iter +
 * the tolerance / strict-convergence tests per iteration) for callers that poll vidseg_kmeans_flags_async themselves.
 * Used by the run-sharded multi-GPU fit: the n_init initialisations of KMeans(n_init=10) (feature_extraction.py:52)
 * are independent, so every rank prepares a workspace for its own subset of the runs and iterates it locally. */
int vidseg_kmeans_lloyd(void* workspace, size_t workspace_bytes, int iterations, void* stream);

/* Whole single-GPU fit+predict: prepare, seed, Lloyd (polling the convergence
 * flags every few iterations), inertia, select, predict.
 *   first_idx_host int32 [R], rand_host double [R, K-1, T]
 *   labels_out int32 [N] (device), centers_out float32 [K, D] (device)
 *   info_host int32[4] = {best run, n_iter of best, max n_iter, kernel launches}
 * Synchronises the stream before returning. */
int vidseg_kmeans_fit_predict(const float* x, int n, int d, int k, int n_init, int n_trials,
                              int max_iter, float tol_rel,
                              const int32_t* first_idx_host, const double* rand_host,
                              int32_t* labels_out, float* centers_out,
                              int32_t* info_host, float* inertia_host,
                              void* workspace, size_t workspace_bytes, void* stream);

/* forget the host-side shape record of a workspace (call before freeing it). */
int vidseg_kmeans_release(void* workspace);

/* ------------------------------------------------------------------------- *
 * R3  correspondence-based mask refinement
 *
 * Replaces scripts/sampling/feature_extraction.py:176-323 (dense nearest-
 * neighbour tracking t -> t+1 with the frame-0 "aux" blend and the per-500-point
 * re-normalisation), :326-364 and :367-461 (signed-jump filter, majority label
 * per trajectory, last-writer-wins write-back).
 *   feats     : float32 [2F, hw, C]  (features of ONE block, uncond rows first)
 *   labels_in : int32   [F, hw]      K-means label maps
 *   traj_out  : int32   [F, hw]      cell index h*W+w of every trajectory per frame
 *   keep_out  : int32   [hw]         1 if the trajectory survives the spatial filter
 *   labels_out: int32   [F, hw]      refined label maps
 * ------------------------------------------------------------------------- */
size_t vidseg_refine_workspace_bytes(int num_frames, int hw, int channels);
int vidseg_refine_masks(const float* feats, const int32_t* labels_in,
                        int num_frames, int height, int width, int channels,
                        int32_t* traj_out, int32_t* keep_out, int32_t* labels_out,
                        void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------- *
 * match_gt_mask mode
 *
 * Replaces scripts/sampling/feature_extraction.py:589-594 (every K-means label of frame 0 takes the most frequent
 * ground-truth label of its cells; ties -> smallest label) and :606-612
 * (sklearn KNeighborsClassifier(n_neighbors=4).fit(ref_feature_map, ref_mask).predict(tokens): brute force,
 * float64 squared distances, neighbours ordered by (distance, index), uniform vote, smallest label on ties).
 *   fake_labels, gt_labels: int32 [n]; gt values must lie in [0, 1024); err_flag int32 (device): 1 = out of range.
 *   ref fp32 [n_ref, D], ref_labels int32 [n_ref], query fp32 [n_query, D]; D % 8 == 0; k <= 8.
 *   err_flag: 2 = more than 64 references tie inside the filter band of the k-th neighbour (duplicated rows).
 * vidseg_knn_predict reads one scalar back (max reference norm) and therefore synchronises the stream once.
 * ------------------------------------------------------------------------- */
int vidseg_majority_map(const int32_t* fake_labels, const int32_t* gt_labels, int n, int num_fake,
                        int32_t* ref_labels, int32_t* err_flag, void* stream);
size_t vidseg_knn_workspace_bytes(int n_ref, int n_query, int d);
int vidseg_knn_predict(const float* ref, const int32_t* ref_labels, int n_ref, const float* query, int n_query, int d,
                       int k, int32_t* labels_out, int32_t* err_flag, void* workspace, size_t workspace_bytes,
                       void* stream);

/* -------------------------------------------------------------------------
 * R7  Sampler step
 * One Euler step of the EDM sampler on the latent, fused: replaces the elementwise chain between two UNet calls of
 * EDMSampler.sampler_step (sgm/modules/diffusionmodules/sampling.py:102-132: Denoiser.forward's
 * net * c_out + input * c_skip, denoiser.py:41-48; the guider's x_u + scale * (x_c - x_u), guiders.py:28-31, 81-90;
 * to_d, sampling_utils.py:34-35; euler_step, sampling.py:92-93) and the latent blending of
 * EulerEDMSampler.__call__ (sampling.py:229-250: x * mask + ori_xt * (1 - mask), mask upsampled nearest-neighbour).
 * Same IEEE fp32 operations in the same order as the reference's eager chain: bit-identical given the same network output.
 *   x [B,C,H,W]; net [G*B,C,H,W] (G = 2 when guided: unconditional half first); c_skip, c_out [G*B]; scale [B]
 *   (guided only); sigma_hat, sigma_next [B]; mask [B, mask_h, mask_w] fp32 or fp64 (mask_is_f64: the blend then runs
 *   in float64 and is rounded once, as torch's type promotion does) with ori_xt [B,C,H,W], or both null; out [B,C,H,W].
 * ------------------------------------------------------------------------- */
int vidseg_sampler_step(const float* x, const float* net, const float* c_skip, const float* c_out, const float* scale,
                        const float* sigma_hat, const float* sigma_next, const void* mask, int mask_is_f64,
                        const float* ori_xt, float* out, int batch, int channels, int height, int width, int mask_h,
                        int mask_w, int guided, void* stream);

/* ------------------------------------------------------------------------- *
 * A1-A8  UNet forward on the tcgen05 tensor cores
 *
 * Split operands: an fp32 tensor x is carried as hi = fp16(x) and lo = fp16(x - hi), x ~= hi + lo (22 significant
 * bits for |x| >= 0.25, absolute error <= 3e-8 below).  Every product is three tensor-core MMAs
 * (lo.hi + hi.lo + hi.hi) accumulated in one fp32 TMEM accumulator.  Tensors of systematically small values are
 * split pre-scaled by a power of two (weights: 2^8) and the GEMM multiplies its accumulator by the exact inverse
 * (acc_scale).  Activations are channels-last ([B, H, W, C] / [B, N, C]).
 * ------------------------------------------------------------------------- */
/* Operand policy of the tensor-core GEMMs (process-wide; set it before the first forward, weights are converted once).
 *   0  every operand is an fp16 pair (hi, lo): three kind::f16 MMAs per product, 22 significant bits;
 *   1  (default) operands whose row length is a multiple of 64 are "packed8": hi = fp16(x) plus, in the memory of `lo`
 *      (same size), per 64-element block 64 bytes of e5m2((x - hi) * 16) followed by 64 bytes of e4m3(x); weights,
 *      which already carry 2^8, use e5m2(w - hi) and e4m3(w / 16).  The two correction products are 2^-11 of the
 *      result and need ~4 significant bits, so they run as kind::f8f6f4 MMAs at twice the fp16 rate: a product costs
 *      2 fp16-MMA units instead of 3 at 2^-14.5 relative precision (1.6e-5 rms per GEMM against fp64; the bar of the
 *      path is 1e-3 on the stashed features).  Attention q / k / v, the K-means filter and operands with other row
 *      lengths stay fp16 pairs.  Every kernel that produces an operand follows the same rule, so callers only ever
 *      pass the (hi, lo) pointer pair around. */
int vidseg_set_operand_mode(int mode);
int vidseg_get_operand_mode(void);

/* fp32 [rows, cols] times `scale` -> operand in the policy's format for rows of `cols` elements; is_weight selects the
 * weight scales of the packed8 format (scale is then the 2^8 weight pre-scale). */
int vidseg_split_rows(const float* x, void* hi, void* lo, long long rows, int cols, float scale, int is_weight,
                      void* stream);

/* elementwise split of n fp32 values times `scale` into an fp16 PAIR, whatever the policy (x 16-byte aligned). */
int vidseg_split_f16(const float* x, void* hi, void* lo, long long n, float scale, void* stream);

/* out[M,N] = A[M,K] . W[N,K]^T (+ bias[N]) (+ residual[M,N]).  Replaces the nn.Linear call sites
 * sgm/modules/attention.py:308,315,317 (to_q/k/v; out_f32 is what the reference stashes as self.q / self.k,
 * :330-331), :364 (to_out), :95 (GEGLU.proj), :110-112 (ff.net[2]), :903,:923 (proj_in / proj_out) and the
 * time-embedding Linears (openaimodel.py:552-556, :281-287).  a_*: [M,K]; w_*: [N,K] (nn.Linear weight layout);
 * out_f32: fp32 [M,N] or NULL; out_hi/out_lo: split of the result or NULL.  N % 4 == 0 (8 with a split output),
 * K % 8 == 0. */
int vidseg_gemm_split(const void* a_hi, const void* a_lo, const void* w_hi, const void* w_lo,
                      const float* bias, const float* residual, float* out_f32,
                      void* out_hi, void* out_lo, int m, int n, int k, float acc_scale, void* stream);

/* GEGLU.proj + gate in one kernel (sgm/modules/attention.py:89-96: x, gate = proj(x).chunk(2, -1); x * gelu(gate)).
 *   w_*: [2*D, K] weight whose rows were permuted so that every 64-row group is 32 value rows followed by the 32 gate
 *        rows of the same output features: row 64*j + i = proj.weight[32*j + i], row 64*j + 32 + i = proj.weight[D + 32*j + i];
 *   bias [2*D] permuted the same way (or NULL); out_*: operand [M, D] in the policy's format.
 * The fp32 [M, 2*D] intermediate is never written (1.17 GB per first-level FeedForward at config 2). */
int vidseg_gemm_geglu_split(const void* a_hi, const void* a_lo, const void* w_hi, const void* w_lo, const float* bias,
                            void* out_hi, void* out_lo, int m, int d, int k, float acc_scale, void* stream);

/* vidseg_gemm_split with the epilogue terms of the SVD temporal layers:
 *   row_bias [M / rows_per_bias, N]: one bias row per group of rows_per_bias consecutive output rows -- the
 *     frame-position embedding 'x_mix = x + emb' (sgm/modules/video_attention.py:417-427, 452-453; group = the
 *     h*w tokens of a frame) and the single-token cross-attention output, which is one vector per sample;
 *   blend, blend_alpha [M / rows_per_alpha]: out = a * blend + (1 - a) * out -- AlphaBlender
 *     (sgm/modules/diffusionmodules/util.py:368-391) as used by SpatialVideoTransformer.time_mixer
 *     (video_attention.py:472-476).
 *   row_scalar [M]: one value per output row added to all of its columns -- the mask modulation
 *     'attn_out[i + num_masks] += lambda * mask[:, None]' of sgm/modules/attention.py:646-663, 697-719, 730-752.
 * Order: acc * acc_scale + bias + row_bias + row_scalar + residual, then the blend.  NULL disables a term.
 * out_pair16 != 0 forces the split output into the fp16-pair format whatever the operand policy says: the q / k / v
 * projections, whose consumer is vidseg_attention_split. */
int vidseg_gemm_split_ex(const void* a_hi, const void* a_lo, const void* w_hi, const void* w_lo,
                         const float* bias, const float* residual, const float* row_bias,
                         long long rows_per_bias, const float* blend, const float* blend_alpha,
                         long long rows_per_alpha, const float* row_scalar, float* out_f32, void* out_hi,
                         void* out_lo, int out_pair16, int m, int n, int k, float acc_scale, void* stream);

/* Up to three projections of the SAME activation as one GEMM: to_q | to_k | to_v of a self-attention layer
 * (sgm/modules/attention.py:308-317: q = self.to_q(x); k = self.to_k(context); v = self.to_v(context) with
 * context = x).  w_*: the nseg weights stacked along N, [nseg * seg_n, K] split; segment s goes to out_f32[s]
 * (fp32 [M, seg_n], NULL = not wanted) and / or out_hi[s] / out_lo[s] (operand [M, seg_n], NULL = not wanted); the
 * three arrays are HOST arrays of nseg device pointers.  The A operand is read once instead of nseg times.
 * seg_n % 128 == 0 or seg_n % 160 == 0 (an N tile must not straddle two outputs).  out_pair16 as in _ex. */
int vidseg_gemm_split_seg(const void* a_hi, const void* a_lo, const void* w_hi, const void* w_lo, int nseg, int seg_n,
                          float* const* out_f32, void* const* out_hi, void* const* out_lo, int out_pair16, int m,
                          int k, float acc_scale, void* stream);

/* nn.Conv2d as an implicit GEMM (no im2col: the taps are shifted TMA boxes, zero padding is the TMA out-of-bounds
 * fill).  Replaces the 3x3 / 1x1 convolutions of ResBlock (openaimodel.py:267-315), Downsample (:202-209, stride 2),
 * Upsample (:145-147), the input conv (:587-593) and the output conv (:825-829).
 *   x_*: [B, H, W, Cin] split;  w_*: [Cout, ksize*ksize*Cin] split, tap-major (ky, kx, cin) = weight.permute(0,2,3,1);
 *   bias [Cout] or NULL; chan_bias [B, Cout] or NULL (ResBlock time embedding, :352-362); residual [B,Ho,Wo,Cout] or NULL;
 *   out_f32 [B, Ho, Wo, Cout] and/or its split.  ksize 1|3 (padding ksize/2), stride 1|2 (3x3 only; Ho = H/2).
 *   Cin % 8 == 0 (64 for stride 2), Cout % 4 == 0. */
int vidseg_conv2d_split(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, const float* bias,
                        const float* chan_bias, const float* residual, float* out_f32, void* out_hi, void* out_lo,
                        int batch, int height, int width, int cin, int cout, int ksize, int stride, float acc_scale,
                        void* stream);

/* Conv3d with a (3,1,1) kernel, padding (1,0,0), over the frame axis of a clip: the time_stack ResBlock of
 * VideoResBlock (sgm/modules/diffusionmodules/video_model.py:45-58, 75-80).  Implicit GEMM like vidseg_conv2d_split;
 * the three taps are the same pixel patch shifted by -1 / 0 / +1 frames.
 *   x_*: [V, T, HW, Cin] split (the '(b t) h w c' channels-last memory of the 2-D layers, no rearrange);
 *   w_*: [Cout, 3*Cin] split, tap-major = weight[:, :, :, 0, 0].permute(0, 2, 1);
 *   frame_bias [V*T, Cout] or NULL: the per-frame time embedding (exchange_temb_dims, openaimodel.py:364-366);
 *   residual [V, T, HW, Cout] or NULL; blend / blend_alpha [V*T]: AlphaBlender 'b t -> b 1 t 1 1'
 *   (video_model.py:59-63, 81-85): out = a * blend + (1 - a) * out. */
int vidseg_conv_temporal_split(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo,
                               const float* bias, const float* frame_bias, const float* residual,
                               const float* blend, const float* blend_alpha, float* out_f32, void* out_hi,
                               void* out_lo, int videos, int frames, int hw, int cin, int cout, float acc_scale,
                               void* stream);

/* softmax(q k^T * scale) v per (batch, head), head dim 64, no mask: replaces
 * F.scaled_dot_product_attention at sgm/modules/attention.py:352-356 (xformers :473-485).
 * q_*: [B, Nq, heads*64] split; k_*, v_*: [B, Nk, heads*64] split (the pre-head-split layout the
 * reference stashes); out_f32 fp32 [B, Nq, heads*64] and/or its split out_hi/out_lo. */
int vidseg_attention_split(const void* q_hi, const void* q_lo, const void* k_hi, const void* k_lo,
                           const void* v_hi, const void* v_lo, float* out_f32, void* out_hi,
                           void* out_lo, int batch, int heads, int nq, int nk, float scale,
                           void* stream);

/* Self-attention over the T <= 32 frames of every spatial site: VideoTransformerBlock.attn1 on the '(b s) t c'
 * layout (sgm/modules/video_attention.py:152, 195; CrossAttention.forward attention.py:286-364).  q, k, v, out are
 * fp32 / split tensors in the frame-major layout [V, T, S, heads*64] of the spatial layers; the reference's
 * '(b t) s c -> (b s) t c' rearrange is index arithmetic inside the kernel.  HBM-bound (16 bytes per element). */
int vidseg_temporal_attention(const float* q, const float* k, const float* v, float* out_f32, void* out_hi,
                              void* out_lo, int videos, int frames, int sites, int heads, float scale,
                              void* stream);

/* nn.LayerNorm over the last dim (attention.py:567-569) -> split operand.  x fp32 [rows, C]; C % 4 == 0, <= 2048. */
int vidseg_layernorm_split(const float* x, const float* gamma, const float* beta, float eps, void* out_hi,
                           void* out_lo, long long rows, int channels, void* stream);

/* LayerNorm(x + row_bias[row / rows_per_bias]) -> split operand: the norms that follow a per-frame / per-sample
 * broadcast add in the temporal layers (norm_in after 'x + emb', video_attention.py:155-157, 452; norm3 after a
 * single-token cross-attention).  row_bias NULL = vidseg_layernorm_split. */
int vidseg_layernorm_bias_split(const float* x, const float* row_bias, long long rows_per_bias, const float* gamma,
                                const float* beta, float eps, void* out_hi, void* out_lo, long long rows,
                                int channels, void* stream);

/* GEGLU gate (attention.py:95-96): h fp32 [rows, 2*D] = (value | gate) -> split(value * gelu_erf(gate)) [rows, D]. */
int vidseg_geglu_split(const float* h, void* out_hi, void* out_lo, long long rows, int d, void* stream);

/* GroupNorm (+ SiLU) over the channel concatenation [x1 (c1) | x2 (c2)] of channels-last fp32 tensors [B, HW, c]
 * -> split operand [B, HW, c1+c2]; raw_hi/raw_lo (optional) receive the split of the un-normalised concatenation
 * (operand of the ResBlock's 1x1 skip convolution).  Replaces GroupNorm32+SiLU (openaimodel.py:267-271, 300-303,
 * 825-827; util.py:276-278), Normalize (attention.py:127-130) and th.cat([h, hs.pop()], 1) (openaimodel.py:911).
 * x2 may be NULL (c2 = 0).  Deterministic (fixed summation order). */
size_t vidseg_groupnorm_workspace_bytes(int batch, int groups);
int vidseg_groupnorm_split(const float* x1, int c1, const float* x2, int c2, const float* gamma, const float* beta,
                           float eps, int groups, int silu, void* out_hi, void* out_lo, void* raw_hi, void* raw_lo,
                           int batch, int hw, void* workspace, size_t workspace_bytes, void* stream);

/* nearest x2 upsampling (openaimodel.py:153) -> split operand.  x fp32 [B, H, W, C] -> [B, 2H, 2W, C]. */
int vidseg_upsample2x_split(const float* x, void* out_hi, void* out_lo, int batch, int h, int w, int c, void* stream);

/* Live per-kernel timing for bench.py's roofline object: while enabled, every launch of the library is
 * bracketed by CUDA events on the stream it is launched on.  Families: 0 other, 1 tcgen05 GEMM (linear layers),
 * 2 tcgen05 attention, 3 tcgen05 convolution, 4 aggregate/normalise, 5 K-means, 6 refine, 7 elementwise/norm.
 * vidseg_profile_enable(1) resets the accumulators; vidseg_profile_read synchronises on the recorded events and
 * returns total milliseconds, launch count and total algorithmic work (FLOPs for 1-3, bytes otherwise; 0 where
 * the launch site does not state it). */
int vidseg_profile_enable(int on);
int vidseg_profile_read(int family, double* ms_total_host, long long* launches_host, double* work_total_host);

/* number of kernel launches issued by this library since load (for bench.py's
 * gpu_launches accounting). */
long long vidseg_launch_count(void);

/* ------------------------------------------------------------------------- *
 * F3  first stage (VAE) around the UNet: sgm/modules/diffusionmodules/model.py:487-748,
 *     sgm/modules/autoencoding/temporal_ae.py:18-110, 293-349.  It runs on the convolution / GroupNorm / GEMM entry
 *     points above plus the two below.
 * ------------------------------------------------------------------------- */
/* Downsample (model.py:77-94): F.pad(x, (0,1,0,1)) then a 3x3 stride-2 convolution without padding.  Operands and
 * outputs as vidseg_conv2d_split; H, W even, Cin % 64 == 0. */
int vidseg_conv2d_down_pad_after_split(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo,
                                       const float* bias, float* out_f32, void* out_hi, void* out_lo,
                                       int batch, int height, int width, int cin, int cout,
                                       float acc_scale, void* stream);
/* softmax over the rows of logits * scale (AttnBlock, model.py:183-189: one head over all H*W sites), written as the
 * split operand [rows, cols] of the following p.v GEMM. */
int vidseg_softmax_rows_split(const float* logits, float scale, void* out_hi, void* out_lo,
                              long long rows, int cols, void* stream);

/* ------------------------------------------------------------------------- *
 * F4  segmentation-map post-process
 *
 * Replaces scripts/sampling/process_output.py: compute_difference :8-28 (uint8 wrap-around difference of the decoded
 * +lambda / -lambda frames, squared in uint8, colour sum, sqrt, cv2.GaussianBlur 5x5 sigma 3 in float64, Pillow
 * float64 -> "L"), the JPEG quality-75 save (:19) / load (:122) round trip of the stored difference image,
 * filter_difference_map :30-38 (Pillow LANCZOS resize of the per-label 0/255 K-means masks) and the normalise +
 * arg-max of get_seg_map_main :120-160.  Every step is bit-exact with the library routine the reference calls.
 *   frames_pos / frames_neg : uint8 [images, H, W, 3]  (images = masks * frames, mask-major)
 *   diff_l   : uint8 [images, H, W]  the image the reference saves as difference_map/original_map/.../{frame}.jpg
 *   vis_l    : uint8 [images, H, W] or NULL: .../vis_map/.../{frame}.jpg (difference / max * 255)
 *   back_l   : uint8 [images, H, W]  diff_l after the JPEG round trip;  back_max int32 [images] its per-image maximum
 *   blur_max : double [images] scratch (maximum of the blurred float64 map)
 * ------------------------------------------------------------------------- */
int vidseg_segmap_difference(const uint8_t* frames_pos, const uint8_t* frames_neg, int images,
                             int height, int width, uint8_t* diff_l, uint8_t* vis_l,
                             uint8_t* back_l, int32_t* back_max, double* blur_max, void* stream);

/* Pillow Image.resize((width, height), LANCZOS) of every (label, frame) mask 255 * (label_maps[f] == unique_labels[m]):
 * label_maps int32 [frames, h, w] -> out uint8 [masks, frames, height, width]; tmp uint8 [masks, frames, h, width].
 * The coefficient windows are Pillow's (precompute_coeffs + normalize_coeffs_8bpc, data independent, computed by the
 * host mirror): bounds int32 [size, 2] = (first input index, count), coeffs int32 [size, ksize] with 22 fractional bits. */
int vidseg_lanczos_masks(const int32_t* label_maps, const int32_t* unique_labels, int masks, int frames,
                         int h, int w, int height, int width,
                         const int32_t* h_bounds, const int32_t* h_coeffs, int h_ksize,
                         const int32_t* v_bounds, const int32_t* v_coeffs, int v_ksize,
                         uint8_t* tmp, uint8_t* out, void* stream);

/* back_l / (back_max + 1e-5), optionally filtered by d*m + filter_s*d*(1-m) with m = mask_resized / 255 (NULL = no
 * filter), arg-max over masks (first maximum) -> seg_raw uint8 [frames, H, W] = unique_labels[argmax] (the
 * segmentation_map_raw PNG content); seg_index int32 [frames, H, W] or NULL = the arg-max itself (colour-map index). */
int vidseg_segmap_argmax(const uint8_t* back_l, const int32_t* back_max, int masks, int frames,
                         int height, int width, const uint8_t* mask_resized, double filter_s,
                         const int32_t* unique_labels, uint8_t* seg_raw, int32_t* seg_index, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VIDSEG_B200_H_ */
