/*
 * abnet3_b200.h -- C ABI of the B200-native ABnet3 hot path.
 *
 * One shared library (abnet3_b200/libabnet3_b200.so), plain C signatures:
 * device pointers, sizes and a CUDA stream; no torch / C++ types.  Every
 * entry point ENQUEUES work on the caller's stream and returns immediately
 * (0 = ABN_OK, otherwise an errno-style code; abn_last_error() gives the
 * text).  Nothing allocates behind the caller's back: outputs and workspaces
 * are caller-owned device buffers.  Re-entrant per stream; one host thread
 * per GPU (one process per GPU under torchrun).
 *
 * The reference (bootphon/abnet3) has no FFI: its boundary for this path is
 * the Python surface cited at each entry point below (paths relative to
 * /root/reference).  INTEGRATION.md shows the ctypes stub a maintainer of the
 * reference would add.
 *
 * Token / pair encoding used throughout
 * -------------------------------------
 *   feat      [n_rows, dim] float32, row-major: every feature file of the
 *             corpus concatenated (abnet3/utils.py:211-217 loads them all,
 *             :122-125 forces float32).  dim % 4 == 0, 16-byte aligned base.
 *   pair_tok  [n_pairs, 4] int32: (row_start_1, n_frames_1, row_start_2,
 *             n_frames_2) -- the two tokens of a pair as row ranges of feat,
 *             i.e. what Features_Accessor.get / get_between_frames slice
 *             (abnet3/utils.py:128-145).
 */
#ifndef ABNET3_B200_H
#define ABNET3_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ABN_OK       0
#define ABN_EIO      5   /* CUDA runtime / launch error                  */
#define ABN_ENOMEM  12   /* workspace too small                          */
#define ABN_EINVAL  22   /* bad argument                                 */
#define ABN_ERANGE  34   /* size outside what the kernels support        */
#define ABN_ENOSYS  38   /* device is not sm_100 (no fallback by design) */

/* longest token (frames) abn_align_pairs / abn_cosine_distance / abn_dtw_from_dist
 * accept; `max_frames` arguments are an upper bound on every n1, n2 of the call
 * (it sizes shared memory; a pair exceeding it comes back with valid = 0).
 * Tokens up to 96 frames take the fused single-tile kernels; longer ones a tiled one. */
#define ABN_MAX_TOKEN_FRAMES 512

typedef void *abn_stream_t; /* a cudaStream_t */

#if defined(__GNUC__)
#define ABN_API __attribute__((visibility("default")))
#else
#define ABN_API
#endif

/* library version (major*10000 + minor*100 + patch) and last error text of
 * the calling thread */
ABN_API int abn_version(void);
ABN_API const char *abn_last_error(void);

/* Device facts for launch sizing / bench reporting.  Any pointer may be NULL. */
ABN_API int abn_device_info(int *sm_count, int *cc_major, int *cc_minor,
                    size_t *smem_optin_bytes);

/* `stack` argument of abn_align_pairs / abn_cosine_distance.
 * 0: generic kernels, any table.  7: the caller vouches that `feat` (dim 280) is the
 * 7-frame stack of 40-wide frames abnet3/features.py:135-159 builds -- row t =
 * [x[t-3] .. x[t+3]], zeros outside the file -- so consecutive rows of a file overlap
 * by 240 columns.  The kernels then read each 40-wide frame once and rebuild the
 * 280-long dot products as 7-tap diagonal sums: 7x fewer FLOPs and bytes, and the
 * SAME BITS as the generic kernels (which accumulate per 40-wide block in that order).
 * abn_stack_violations counts the rows r (not flagged in last_row_of_file, NULL = none)
 * whose columns [dim/stack, dim) differ from row r+1's columns [0, dim - dim/stack);
 * `count` (device, caller-zeroed) == 0 means the table qualifies. */
ABN_API int abn_stack_violations(const float *feat, int64_t n_rows, int dim, int stack,
                                 const uint8_t *last_row_of_file, unsigned long long *count,
                                 abn_stream_t stream);
/* Host -> device upload of a table the caller vouches to be such a stack: only the middle
 * block of every row crosses PCIe (a strided 2-D copy from the HOST pointer feat_host,
 * ideally pinned: `stack` x fewer bytes), then a kernel rebuilds the other blocks from the
 * neighbouring rows' middle blocks, with zeros across file edges (last_row_of_file: device,
 * [n_rows] uint8, NULL = one file).  feat_dev [n_rows, dim] ends up identical to feat_host. */
ABN_API int abn_stack_upload(float *feat_dev, const float *feat_host, int64_t n_rows, int dim,
                             int stack, const uint8_t *last_row_of_file, abn_stream_t stream);
/* First-class UN-STACKED input: `frames` [n_rows, f] float32 (DEVICE: the caller
 * uploads the frames as they are, one contiguous copy, stack x fewer bytes than the stacked
 * table) holds the f-wide frames themselves -- what abnet3/features.py:135-159 stacks -- and
 * last_row_of_file (device, [n_rows] uint8, NULL = one file) marks the file edges.  Builds
 * feat_dev [n_rows, stack * f]: row t = [x[t - stack/2] .. x[t + stack/2]], zeros outside the
 * file, exactly the reference's stack_fbanks.  Nothing has to be vouched for. */
ABN_API int abn_stack_from_frames(float *feat_dev, const float *frames, int64_t n_rows, int f,
                                  int stack, const uint8_t *last_row_of_file, abn_stream_t stream);

/* Compact form of the alignment result for the trip back to the host: the path of pair p
 * (dense table idx1 / idx2 [dst_off[p] .. dst_off[p] + path_len[p]) from abn_compact_paths)
 * as its step DIRECTIONS, 2 bits each (0 = diagonal (+1, +1), 1 = up (+1, +0), 2 = left
 * (+0, +1)), four per byte, path_len[p] - 1 of them starting at byte dir_off[p] of `dirs`
 * (dir_off[p + 1] - dir_off[p] = ceil((path_len[p] - 1) / 4), caller-computed prefix sums).
 * Every path starts at (row1, row2) of the pair, so directions + pair_tok restore the index
 * pairs (abnet3_b200.utils.decode_directions): 32 x fewer bytes than two int32 per step. */
ABN_API int abn_pack_directions(const int32_t *idx1, const int32_t *idx2, const int64_t *dst_off,
                                const int32_t *path_len, int n_pairs, const int64_t *dir_off,
                                uint8_t *dirs, abn_stream_t stream);

/* Device workspace for abn_align_pairs / abn_cosine_distance over n_pairs pairs.
 * Both bucket the pairs into token-length classes on the device and keep the
 * class-sorted pair order there.  abn_align_pairs additionally hands every pair's
 * distance matrix (float32, at most min(max_frames, 96)^2 cells) from the distance
 * kernels to the DTW kernel through this workspace; it works through the sorted pair
 * list in `rounds` windows, so a smaller workspace only means more, shorter rounds.
 * Returns the bytes for the given number of rounds (rounds = 1: every pair has its own
 * slot).  abn_align_pairs accepts ANY size >= abn_align_workspace_bytes(n_pairs,
 * max_frames, n_pairs) and derives the round count from what it is given;
 * abn_cosine_distance needs abn_align_workspace_bytes(n_pairs, 1, n_pairs). */
ABN_API size_t abn_align_workspace_bytes(int n_pairs, int max_frames, int rounds);
/* Number of kernels one abn_align_pairs call with this workspace enqueues (bucketing +
 * per round: one distance kernel per size class and one DTW kernel per class row). */
ABN_API int abn_align_launches(int n_pairs, int max_frames, int stack, size_t workspace_bytes);

/* ------------------------------------------------------------------------
 * (1) Batched cosine frame distance.
 * Replaces abnet3/utils.py:40-60 `cosine_distance(x, y)`, one call per pair.
 *   dist_off [n_pairs+1] int64: pair p's n1*n2 matrix (row-major, ld = n2)
 *            starts at dist[dist_off[p]].
 *   dist     float32.  The reference computes in float32 and widens to
 *            float64 at the end (utils.py:49-53), so float32 storage holds
 *            exactly the values the reference returns.
 *   valid    [n_pairs] uint8, 0 when the matrix holds a NaN / negative entry
 *            (the reference's `assert np.all(d >= 0)`, utils.py:59).
 * ---------------------------------------------------------------------- */
ABN_API int abn_cosine_distance(const float *feat, int64_t n_rows, int dim,
                        const int32_t *pair_tok, int n_pairs, int max_frames, int stack,
                        const int64_t *dist_off, float *dist, uint8_t *valid,
                        void *workspace, size_t workspace_bytes,
                        abn_stream_t stream);

/* ------------------------------------------------------------------------
 * (2) Batched DTW + traceback on GIVEN float64 distance matrices.
 * Replaces the external call at abnet3/utils.py:149-151
 *   `DTW(feat1, feat2, return_alignment=True, dist_array=distance_array)`.
 *   shape    [n_pairs, 2] int32 (n1, n2); dist/dist_off as above but float64.
 *   path_off [n_pairs+1] int64 with path_off[p+1]-path_off[p] >= n1+n2-1.
 *   path1/2  LOCAL frame indices, forward order, at path_off[p] .. +path_len.
 *   cost     C[n1-1, n2-1]; valid as above.
 * Tie rule: diagonal, then i-1, then j-1 (oracle/dtw_oracle.c).
 * Test hook for the bit-exact mode: matrices up to 96 x 96.
 * ---------------------------------------------------------------------- */
ABN_API int abn_dtw_from_dist(const double *dist, const int64_t *dist_off,
                      const int32_t *shape, int n_pairs, int max_frames,
                      const int64_t *path_off, int32_t *path1, int32_t *path2,
                      int32_t *path_len, double *cost, uint8_t *valid,
                      abn_stream_t stream);

/* ------------------------------------------------------------------------
 * (1)+(2) fused: align every 'same' pair of a pair list.
 * Replaces abnet3/utils.py:147-153 `get_dtw_alignment` as called per pair by
 * abnet3/dataloader.py:183-206 and :642-653.  Per size class: a distance kernel
 * (CTA per pair, shared-memory staged) leaves the float32 matrix in the workspace in
 * anti-diagonal order, then a DTW kernel (one warp per pair) sweeps it; accumulated
 * costs and traceback directions never leave the SM.  Tokens longer than 96 frames
 * (up to 512) take a tiled kernel that keeps everything on chip.
 *   idx1/idx2 GLOBAL row ids into feat (row_start + local path index), i.e.
 *             the rows `feat1[path1, :]`, `feat2[path2, :]` gather
 *             (dataloader.py:204-205), at path_off[p] .. +path_len[p].
 *   valid[p] == 0 reproduces "exception -> pair dropped"
 *             (dataloader.py:188-191); then path_len[p] = 0.
 * ---------------------------------------------------------------------- */
ABN_API int abn_align_pairs(const float *feat, int64_t n_rows, int dim,
                    const int32_t *pair_tok, int n_pairs, int max_frames, int stack,
                    const int64_t *path_off, int32_t *idx1, int32_t *idx2,
                    int32_t *path_len, double *cost, uint8_t *valid,
                    void *workspace, size_t workspace_bytes,
                    abn_stream_t stream);

/* ------------------------------------------------------------------------
 * Frame-index pairs of 'diff' pairs (no DTW).
 * Replaces abnet3/dataloader.py:208-231 (and :655-666):
 *   stretch == 0: both tokens truncated to min(n1,n2) leading frames;
 *   stretch != 0: `align_different_words` -- the longer token in X1, the
 *                 shorter one stretched by rint(linspace(0, min-1, max)) in X2.
 *   out_off [n_pairs+1] int64: where pair p's rows go; rows written =
 *           min(n1,n2) (stretch == 0) or max(n1,n2).
 * ---------------------------------------------------------------------- */
ABN_API int abn_diff_pairs(const int32_t *pair_tok, int n_pairs, int stretch,
                   const int64_t *out_off, int32_t *idx1, int32_t *idx2,
                   abn_stream_t stream);

/* ------------------------------------------------------------------------
 * Compact the per-pair path slots written by abn_align_pairs into a dense
 * frame-pair table (what FramesDataLoader.load_all_frames builds,
 * dataloader.py:642-653): for every pair with path_len > 0, copy its
 * path_len entries from (src_off[p]) to (dst_off[p]).
 * ---------------------------------------------------------------------- */
ABN_API int abn_compact_paths(const int32_t *src1, const int32_t *src2,
                      const int64_t *src_off, const int64_t *dst_off,
                      const int32_t *path_len, int n_pairs, int32_t *dst1,
                      int32_t *dst2, abn_stream_t stream);

/* One int64 from device memory into PINNED (device-mapped) host memory by a kernel store, on
 * `stream`: how a host-side pipeline learns a data-dependent size (the total path length of a
 * chunk) while a bulk device -> host copy occupies the copy engine -- a cudaMemcpy of the scalar
 * would queue behind it and stall the next chunk's launches. */
ABN_API int abn_store_scalar64(const int64_t *src, int64_t *dst_host_mapped, abn_stream_t stream);

/* ------------------------------------------------------------------------
 * Batch generation: gather feature rows for a batch of frame pairs.
 * Replaces the row gathers of abnet3/dataloader.py:204-205, :250-255 and
 * FramesDataLoader.load_batch (:673-684).
 *   sel   [n] int64 or NULL: positions in idx1/idx2/y_in (the permutation /
 *         batch slice); NULL = 0..n-1.
 *   x1,x2 [n, dim] float32 out; y_out [n] float32 out (NULL to skip);
 *   y_in  [*] int8 labels (+1 / -1) or NULL.
 * ---------------------------------------------------------------------- */
ABN_API int abn_gather_batch(const float *feat, int dim, const int32_t *idx1,
                     const int32_t *idx2, const int8_t *y_in,
                     const int64_t *sel, int64_t n, float *x1, float *x2,
                     float *y_out, abn_stream_t stream);

/* ------------------------------------------------------------------------
 * (4) Fused pair loss + gradient.
 * Replaces abnet3/loss.py:46-67 (coscos2) and :85-105 (cosmargin) together
 * with the autograd backward the trainer runs (abnet3/trainer.py:237-240).
 *   kind   0 = coscos2, 1 = cosmargin(margin)
 *   y      [n] float32 labels (+1 same, -1 diff, anything else: raw cosine)
 *   scale  multiplies loss and gradients (1/n for avg=True, the multitask
 *          weight, ...)
 *   loss   [1] float32, ACCUMULATED into (caller zeroes it)
 *   de1/2  [n, dim] float32 gradients w.r.t. e1 / e2 (NULL: forward only)
 *   ld     row stride (elements) of e1, e2, de1, de2 (0 = dim): the two heads of the
 *          multitask network live side by side in one [n, 2*dim] buffer
 * ---------------------------------------------------------------------- */
ABN_API int abn_pair_loss(const float *e1, const float *e2, const float *y, int64_t n,
                  int dim, int64_t ld, int kind, float margin, float scale, float *loss,
                  float *de1, float *de2, abn_stream_t stream);

/* ------------------------------------------------------------------------
 * Dropout between Linear and activation in TRAINING mode (abnet3/model.py:136-141:
 * `Linear -> Dropout(p) -> [BatchNorm] -> act`, default p = 0.1 at :111): the layer computes
 *   y = act(keep * z / (1 - p)),   z = x W^T + b,   dz = dy * act'(y) * keep / (1 - p).
 * The keep mask is a pure function of (seed, step, layer, row, col) -- a counter-based hash
 * (splitmix64 of the key and the position, 16 bits per element: p is resolved to 1 / 65536) --
 * that every kernel re-evaluates in its epilogue: no mask is stored, and the backward pass
 * sees exactly the forward pass's mask.  `state` is a DEVICE array {seed, step}: kernels read
 * it at run time, so a CUDA-graph replay sees the step the host (or a captured kernel) has
 * advanced.  p == 0 or state == NULL: no dropout (eval mode).
 * abn_dropout_mask writes the mask (1 = kept) of one layer for inspection / replay in tests. */
typedef struct {
    const unsigned long long *state;    /* device {seed, step} */
    float p;                            /* drop probability */
    int layer;                          /* which layer's mask */
} abn_dropout;
ABN_API int abn_dropout_mask(const abn_dropout *drop, int64_t rows, int cols, uint8_t *mask,
                             abn_stream_t stream);

/* ------------------------------------------------------------------------
 * (3) Embedder MLP layers.
 * Replace one `Linear -> Dropout(p=0) -> activation` block of
 * abnet3/model.py:133-170 (forward) and its autograd backward.
 *   act: 0 none, 1 sigmoid, 2 tanh, 3 relu
 *   W [n_out, n_in] row-major float32 (nn.Linear layout), b [n_out].
 *   precision: 0 = fp32 SIMT (the 1e-4 parity path).  The bf16 tcgen05 path needs
 *              bf16 operand buffers and is exposed as abn_gemm_bf16_group /
 *              abn_mlp_forward_fused / abn_mlp_dgrad_fused below.
 * Forward:   y[m, n_out] = act(x[m, n_in] @ W^T + b)
 * Backward:  dz = dy * act'(y);  dx = dz @ W (NULL to skip);
 *            dW += dz^T @ x;  db += colsum(dz)
 *   accumulate: bit 0 set = add into dW / db (else overwrite them);
 *               bit 1 set = add into dx (two heads feeding one trunk)
 * ---------------------------------------------------------------------- */
ABN_API int abn_linear_forward(const float *x, const float *W, const float *b,
                       int64_t m, int n_in, int n_out, int act, int precision,
                       float *y, abn_stream_t stream);
ABN_API int abn_linear_backward(const float *x, const float *W, const float *y,
                        float *dy /* overwritten with dz */, int64_t m,
                        int n_in, int n_out, int act, int precision,
                        int accumulate, float *dx, float *dW, float *db,
                        abn_stream_t stream);
/* The same two calls with dropout on the layer's pre-activation (drop == NULL or p == 0: as
 * above).  Mask rows are row_offset + the row index inside the call. */
ABN_API int abn_linear_forward_drop(const float *x, const float *W, const float *b,
                       int64_t m, int n_in, int n_out, int act, int precision,
                       float *y, const abn_dropout *drop, int64_t row_offset, abn_stream_t stream);
ABN_API int abn_linear_backward_drop(const float *x, const float *W, const float *y,
                        float *dy /* overwritten with dz */, int64_t m,
                        int n_in, int n_out, int act, int precision,
                        int accumulate, float *dx, float *dW, float *db,
                        const abn_dropout *drop, int64_t row_offset, abn_stream_t stream);

/* ------------------------------------------------------------------------
 * (3) tensor-core path, persistent grouped GEMM: up to ABN_GEMM_MAX_GROUP problems
 * C[M,N] = A . B^T in ONE launch (tcgen05 / TMEM / TMA; one persistent CTA per SM walks
 * the 128 x 256 output tiles of all problems, the accumulator double buffered in TMEM).
 * Operands are bf16 arrays in their natural row-major layout, never transposed copies:
 *   a_mn = 0: A is [M, K] (K contiguous);  a_mn = 1: A is [K, M] (M contiguous)
 *   b_mn = 0: B is [N, K] (K contiguous);  b_mn = 1: B is [K, N] (N contiguous)
 * so for one layer of abnet3/model.py:133-170 and its backward:
 *   forward  A = x [rows, n_in],    B = W [n_out, n_in]            a_mn 0, b_mn 0, epilogue 0
 *   dgrad    A = dz [rows, n_out],  B = W [n_out, n_in]            a_mn 0, b_mn 1, epilogue 1
 *   wgrad    A = dz [rows, n_out],  B = x [rows, n_in]  (K = rows) a_mn 1, b_mn 1, epilogue 2
 * Epilogues: 0  out = act(C + bias)            -> bf16 [M, ldo] or fp32 (out_f32)
 *            1  out = C * act'(yprev)          -> the dz of the layer below (yprev = its
 *                                                 forward output, bf16 [M, ld_yprev])
 *            2  out += C (fp32 reds, split_k)  -> out fp32 [M, ldo]
 * Problems of one group are independent unless chained with signal / wait below.
 * ones_col (epilogues 0/1): also write 1.0 into column N of every output row (it must lie
 *   inside the row padding, ldo >= N + 1).  ones_out (epilogue 2): B has such a column of
 *   ones at index N; its products -- the column sums of A, i.e. the bias gradient -- are
 *   added to ones_out[M] instead of `out`.
 * Leading dimensions in elements; bf16 pointers 16-byte aligned, bf16 lds multiples of 8.
 * ---------------------------------------------------------------------- */
#define ABN_GEMM_MAX_GROUP 8
typedef struct {
    const void *A; int64_t lda; int a_mn;
    const void *B; int64_t ldb; int b_mn;
    int M, N, K;
    int epilogue, act, split_k;
    const float *bias;
    void *out; int64_t ldo; int out_f32;
    const void *yprev; int64_t ld_yprev;
    int ones_col;
    float *ones_out;
    /* in-launch dependencies between the problems of one group (NULL: none).  signal
     * [ceil(M / 256)] int32, caller-zeroed: every CTA adds 1 per finished tile of that
     * 256-row block once its output rows are in global memory.  wait / wait_count: the A
     * rows of a block are loaded only once wait[block] >= wait_count -- point `wait` at an
     * earlier problem's `signal` and leave wait_count 0 (= all of that problem's tiles over
     * the block) to chain layers inside one launch (list the problems in dependency order).
     * A reduction problem (epilogue 2: its K dimension runs over the rows, e.g. dW = dz^T x)
     * waits for every block its K range crosses, block by block as the k loop reaches them:
     * the weight gradients of a layer can share the launch of the dgrad chain that produces
     * their dz and fill the gaps its dependencies leave. */
    int32_t *signal;
    const int32_t *wait;
    int wait_count;
} abn_gemm_problem;
ABN_API int abn_gemm_bf16_group(const abn_gemm_problem *problems, int n_problems,
                                abn_stream_t stream);

/* ------------------------------------------------------------------------
 * (3b) the embedder's forward pass in ONE launch with the activations resident on chip:
 * every layer y = act(x W^T + b) of SiameseNetwork.forward_once (abnet3/model.py:179-186)
 * for a 256-row block per CTA pair.  The block's activations stay in shared memory as the
 * next layer's A operand (written there by the epilogue), only the weights stream through
 * TMA; hidden outputs also leave as bf16 rows (the backward pass reads them), the last layer
 * may write fp32.  Same arithmetic, same k order and the same bits as the layers chained
 * through abn_gemm_bf16_group.  Limits: n_in <= 512, n_out (+ ones_col) <= 512 per layer
 * (ABN_EINVAL otherwise: use the grouped GEMM).
 *   x [rows, n_in of layer 0] bf16, ldx elements; W [n_out, n_in] bf16, ldw; out: bf16
 *   [rows, ldo] (ldo % 8 == 0, >= n_out + ones_col) or, last layer only, fp32 [rows, ldo].
 * ---------------------------------------------------------------------- */
#define ABN_MLP_MAX_LAYERS 8
typedef struct {
    const void *W; int64_t ldw;
    const float *bias;          /* NULL: none */
    int n_in, n_out, act;       /* act: ABN_ACT_* */
    void *out; int64_t ldo; int out_f32;
    int ones_col;               /* write 1.0 into column n_out of the bf16 output */
    abn_dropout drop;           /* dropout on this layer's pre-activation (p = 0: none) */
} abn_mlp_layer;
ABN_API int abn_mlp_forward_fused(const void *x, int64_t ldx, int64_t rows,
                                  const abn_mlp_layer *layers, int n_layers, abn_stream_t stream);
/* The same launch with the pair loss FUSED into the last layer's epilogue (abnet3/trainer.py:231-236:
 * forward, then loss): the rows of x are INTERLEAVED -- row 2k = X1[k], row 2k + 1 = X2[k], what
 * abn_gather_step_bf16(interleave = 1) writes -- so both embeddings of a pair meet in one warp;
 * coscos2 (kind 0) / cosmargin (kind 1) value (added to *loss, times scale) and
 * dz = dL/de * act'(e) of the output layer as bf16 rows [rows, ld_dz] (what abn_pair_loss_dz
 * produces) come straight from the accumulators.  Last layer: fp32 output of at most 128 columns,
 * no dropout; the fp32 embeddings are written only if write_embeddings. */
typedef struct {
    const float *y;             /* [rows / 2] labels (+1 / -1 / other, loss.py:59-62) */
    float *loss;                /* accumulated into */
    void *dz; int64_t ld_dz;    /* bf16 [rows, ld_dz] */
    int kind;                   /* 0 coscos2, 1 cosmargin */
    float margin, scale;
    int write_embeddings;
} abn_mlp_loss;
ABN_API int abn_mlp_forward_loss_fused(const void *x, int64_t ldx, int64_t rows,
                                       const abn_mlp_layer *layers, int n_layers,
                                       const abn_mlp_loss *loss, abn_stream_t stream);
/* The input-gradient chain of the backward pass the same way (the autograd of those blocks,
 * abnet3/trainer.py:238): layers listed from the top down, layer l computes
 *   dz_below = (dz . W) * act'(y_below)      W [n_out, n_in] as stored, dz [rows, n_out]
 * with dz of the first listed layer read from dz_top and every dz_below both written to
 * global memory (bf16 rows: the weight-gradient GEMMs read them) and kept on chip as the next
 * layer's A operand; y_below [rows, ld_y] is fetched per warp by TMA.  Limits: n_in, n_out
 * <= 512. */
typedef struct {
    const void *W; int64_t ldw;
    int n_in, n_out;
    int act_below;              /* activation whose output y_below is */
    const void *y_below; int64_t ld_y;
    void *dz_below; int64_t ld_dz;
    abn_dropout drop_below;     /* the dropout of the layer whose output y_below is */
} abn_mlp_dlayer;
ABN_API int abn_mlp_dgrad_fused(const void *dz_top, int64_t ld_top, int64_t rows,
                                const abn_mlp_dlayer *layers, int n_layers, abn_stream_t stream);

/* fp32 [rows, cols] (ld_src) -> bf16 [rows, ld_dst] and/or transposed bf16
 * [cols, ld_T]: operand preparation for abn_gemm_bf16_group (the weights after
 * load_state_dict, an fp32 input batch). */
ABN_API int abn_cast_bf16(const float *src, int64_t rows, int cols, int64_t ld_src, void *dst,
                          int64_t ld_dst, void *dstT, int64_t ld_T, abn_stream_t stream);

/* ------------------------------------------------------------------------
 * Fused optimizer step over one flat parameter bucket.
 * Replaces `optimizer.step()` of abnet3/trainer.py:240 for the optimizers the
 * trainer offers (:68-87) that the configs use: kind 0 = SGD with momentum
 * (torch.optim.SGD, dampening 0), 1 = Adadelta (rho 0.9, eps 1e-6),
 * 2 = Adam (betas 0.9/0.999, eps 1e-8).
 *   grad_scale multiplies the gradient first (1/world_size after a
 *   sum-allreduce).  state0/state1: momentum buffer | (square_avg, acc_delta)
 *   | (exp_avg, exp_avg_sq); step = 1-based step count (Adam bias correction).
 * ---------------------------------------------------------------------- */
ABN_API int abn_optimizer_step(float *param, const float *grad, float *state0,
                       float *state1, int64_t n, int kind, float lr,
                       float momentum, float grad_scale, int64_t step,
                       abn_stream_t stream);

/* ------------------------------------------------------------------------
 * Memory-bound companions of the tensor-core step, fused with the conversion their
 * consumer needs (no fp32 intermediate makes a round trip through HBM).
 *
 * abn_gather_batch_bf16: abn_gather_batch writing bf16 rows -- X1 = rows 0..n-1,
 *   X2 = rows n..2n-1 of xb [2n, ldx] -- i.e. the first layer's A operand.  zero_me
 *   (nullable): zero_words 4-byte words the kernel clears (the step's loss accumulator and
 *   the dependency counters of the chained GEMM launches).
 * abn_pair_loss_dz: abn_pair_loss whose gradient output is dz = dL/de * act'(e) as bf16
 *   rows [n, ld_dz] (act = the output layer's activation; e1/e2 are its outputs): the
 *   operand of the output layer's dgrad / wgrad GEMMs.
 * abn_optimizer_step_fused: abn_optimizer_step over `segments` of the flat bucket; a
 *   segment with bf16 != NULL is a weight matrix [count / n_in, n_in] whose updated values
 *   are also written as bf16 rows of leading dimension ld; zero_grad clears the gradient
 *   after use (the next step's wgrad reductions accumulate into it).
 * ---------------------------------------------------------------------- */
ABN_API int abn_gather_batch_bf16(const float *feat, int dim, const int32_t *idx1,
                                  const int32_t *idx2, const int8_t *y_in, const int64_t *sel,
                                  int64_t n, void *xb, int64_t ldx, float *y_out, void *zero_me,
                                  int zero_words, abn_stream_t stream);
/* abn_gather_step_bf16: abn_gather_batch_bf16 for an EPOCH of fixed-size frame batches over a
 *   device-resident, already shuffled frame-pair table (FramesDataLoader.batch_iterator,
 *   abnet3/dataloader.py:686-739, feeding abnet3/trainer.py:231-243), so that a training step
 *   replays as one CUDA graph with no per-batch host work:
 *   - y2_in / y2_out (nullable): a second label column (y_spk beside y_phn, dataloader.py:753-792);
 *   - cursor (nullable, device int64[2] = {next row, 0}): when sel == NULL the batch is the table
 *     rows cursor[0] .. cursor[0]+n-1 and the kernel advances cursor[0] by n; table_rows > 0:
 *     a batch that would run past row table_rows is not gathered (prefetch of the batch after
 *     the sweep's last one);
 *   - loss_acc (nullable, device double[1]): before zero_me is cleared its word 0 (the previous
 *     step's float loss) is added to loss_acc -- `train_loss += loss.data[0]`, trainer.py:242;
 *   - interleave != 0: X1[k] -> row 2k, X2[k] -> row 2k + 1 of xb (instead of rows k and n + k):
 *     the layout abn_mlp_forward_loss_fused needs. */
ABN_API int abn_gather_step_bf16(const float *feat, int dim, const int32_t *idx1,
                                 const int32_t *idx2, const int8_t *y_in, const int8_t *y2_in,
                                 const int64_t *sel, int64_t *cursor, int64_t table_rows,
                                 int64_t n, void *xb,
                                 int64_t ldx, float *y_out, float *y2_out, void *zero_me,
                                 int zero_words, double *loss_acc, int interleave,
                                 abn_stream_t stream);
ABN_API int abn_pair_loss_dz(const float *e1, const float *e2, const float *y, int64_t n, int dim,
                             int64_t ld, int kind, float margin, float scale, int act, float *loss,
                             void *dz1, void *dz2, int64_t ld_dz, abn_stream_t stream);
/* ... with the output layer's dropout folded into dz (mask rows: row for dz1, row2_offset + row
 * for dz2; mask columns: col_offset + column -- the two multitask heads are one 2d-wide layer) */
ABN_API int abn_pair_loss_dz_drop(const float *e1, const float *e2, const float *y, int64_t n, int dim,
                                  int64_t ld, int kind, float margin, float scale, int act, float *loss,
                                  void *dz1, void *dz2, int64_t ld_dz, const abn_dropout *drop,
                                  int64_t row2_offset, int col_offset, abn_stream_t stream);
#define ABN_MAX_PARAM_SEGMENTS 24
typedef struct {
    int64_t offset, count;      /* range of the flat bucket */
    int64_t ld;                 /* bf16 copy: leading dimension (elements) */
    void *bf16;                 /* NULL: no bf16 copy (biases) */
    int n_in;                   /* bf16 copy: row length of the matrix */
} abn_param_segment;
ABN_API int abn_optimizer_step_fused(float *param, float *grad, float *state0, float *state1,
                                     int kind, float lr, float momentum, float grad_scale,
                                     int64_t step, const abn_param_segment *segments,
                                     int n_segments, int zero_grad, abn_stream_t stream);

/* ------------------------------------------------------------------------
 * Data parallelism without a library collective (new: the reference is single-GPU,
 * abnet3/trainer.py:59-61): the gradient all-reduce fused into the optimizer over NVLink /
 * NVSwitch peer memory, one process per GPU on one box.
 *   abn_ipc_export / abn_ipc_import: CUDA IPC handle (64 bytes) + byte offset of a device
 *     buffer, and the mapping of a peer's buffer into this process (peer access enabled).
 *   abn_dp_peers: every rank's gradient bucket and flag block (ABN_DP_FLAG_WORDS uint64,
 *     zero-initialised once) as mapped in THIS process; entry `rank` is the local one.
 *   abn_dp_optimizer_step: abn_optimizer_step_fused on grad = sum over ranks (read straight
 *     from the peers' buckets once every rank has signalled that its gradients are complete).
 *   abn_dp_grad_reset: clears the local bucket once every peer has finished reading it
 *     (enqueue before the next step's gradient reductions).
 * Every rank must enqueue the same sequence of steps; the flags are device-side step
 * counters, so the sequence may be captured in a CUDA graph and replayed.
 * ---------------------------------------------------------------------- */
#define ABN_DP_MAX_WORLD 8
#define ABN_DP_FLAG_WORDS 32
typedef struct {
    void *grad[ABN_DP_MAX_WORLD];
    void *flags[ABN_DP_MAX_WORLD];
    int rank, world;
} abn_dp_peers;
ABN_API int abn_ipc_export(const void *ptr, unsigned char *handle64, int64_t *offset);
ABN_API int abn_ipc_import(const unsigned char *handle64, int64_t offset, void **ptr);
ABN_API int abn_dp_optimizer_step(float *param, float *state0, float *state1, int kind, float lr,
                                  float momentum, float grad_scale, int64_t step,
                                  const abn_param_segment *segments, int n_segments,
                                  const abn_dp_peers *peers, abn_stream_t stream);
ABN_API int abn_dp_grad_reset(float *grad, int64_t n, const abn_dp_peers *peers,
                              abn_stream_t stream);
/* Two-shot variant in which every transfer is a WRITE over NVLink (posted, no round trips):
 * each rank pushes slice j of its gradient bucket to owner j, the owner reduces in rank
 * order, applies the optimizer to its slice and pushes the updated fp32 parameters into
 * every rank's parameter bucket; then bf16 weight copies and the gradient reset, all in ONE
 * kernel (one resident block per SM).  param[r] / recv[r] / flags[r]: rank r's parameter
 * bucket (n trained floats first), receive buffer (world x slice_cap floats) and flag block
 * as mapped in this process.  n and slice_cap are multiples of 4, world * slice_cap >= n. */
typedef struct {
    void *param[ABN_DP_MAX_WORLD];
    void *recv[ABN_DP_MAX_WORLD];
    void *flags[ABN_DP_MAX_WORLD];
    int rank, world;
    int64_t n, slice_cap;
    /* one_shot != 0: every rank pushes its WHOLE bucket to every peer (receive buffer = 2 x world x
     * slice_cap floats, slice_cap >= n, double buffered by step parity), one flag exchange, then a
     * purely local rank-order sum + update: half the synchronisation latency of the two-shot form
     * for world/2 x its bytes -- the exchange is latency bound at these bucket sizes. */
    int one_shot;
    /* one_shot == 2: two-shot exchange with the flag inside the data (8-byte {value, step} stores,
     * receivers spin on the elements themselves: no fences, no tickets, no flag round trips).
     * recv = uint64 [world * slice_cap] gradient inbox followed by uint64 [n] parameter inbox. */
} abn_dp_push;
ABN_API int abn_dp_push_step(float *grad, float *state0, float *state1, int kind, float lr,
                             float momentum, float grad_scale, int64_t step,
                             const abn_param_segment *segments, int n_segments,
                             const abn_dp_push *peers, abn_stream_t stream);
/* debug hook (tools/dp_trace.py): device buffer [64][8] int64 that block 0 of the push kernels
 * stamps with %globaltimer at its phase boundaries (NULL: off) */
ABN_API int abn_dp_set_trace(long long *buffer);

#ifdef __cplusplus
}
#endif
#endif /* ABNET3_B200_H */
