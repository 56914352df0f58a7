/*
 * gandtr_b200 -- C ABI of the B200-native retrieval hot path (libgandtr_b200.so).
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / C++ types. Every entry point
 * cites the reference interface it replaces (paths relative to the mohwald/gandtr checkout).
 * The reference is pure Python, so its "FFI" is ctypes: gandtr_b200/_lib.py is the binding,
 * INTEGRATION.md shows the stub a maintainer would add on the reference side.
 *
 * Conventions
 *   - every `const T* x` / `T* x` argument named without a `host_` prefix is a DEVICE pointer
 *     owned by the caller (in practice: torch tensors); `host_*` arguments are read on the host
 *     during the call.
 *   - all work is enqueued asynchronously on `stream` (a cudaStream_t passed as void*); nothing
 *     synchronises the device, nothing allocates device memory: scratch comes from the
 *     caller-provided workspace (`ws`, `ws_bytes`; query the size with the matching
 *     gdt_*_workspace_bytes call). Workspaces must be 256-byte aligned.
 *   - return value: GDT_OK (0) or a negative gdt_status. Functions never throw.
 *   - thread-safe for distinct streams; tables are immutable after gdt_init.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns
 *     GDT_ERR_NO_DEVICE / GDT_ERR_NOT_INITIALISED.
 */
#ifndef GANDTR_B200_H
#define GANDTR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GDT_ABI_VERSION 1

typedef enum gdt_status {
    GDT_OK = 0,
    GDT_ERR_INVALID_ARGUMENT = -1,
    GDT_ERR_NOT_INITIALISED = -2,
    GDT_ERR_NO_DEVICE = -3,
    GDT_ERR_WORKSPACE_TOO_SMALL = -4,
    GDT_ERR_CUDA = -5,          /* a CUDA runtime / driver call failed; see gdt_last_cuda_error */
    GDT_ERR_UNSUPPORTED = -6,
    GDT_ERR_CANDIDATE_OVERFLOW = -7 /* score_topk: a query exceeded its candidate capacity (reported
                                       through the device status word, see gdt_score_topk) */
} gdt_status;

/* ---- library ----------------------------------------------------------------------------- */

int gdt_abi_version(void);
const char* gdt_status_string(int status);
/* text of the last CUDA error seen by the calling thread ("" if none) */
const char* gdt_last_cuda_error(void);

/*
 * Upload the immutable tables of the CLAHE path to the *current* CUDA device:
 *   host_rgb2lab_lut : 33*33*33*3 int16, the lattice table of OpenCV's float RGB2Lab
 *                      (color_lab.cpp RGB2LabLUT_s16), layout [r][g][b][L,a,b]
 *                      (shipped as gandtr_b200/data/rgb2lab_lut_s16.bin).
 * The sRGB inverse-gamma spline (1024x4 f32) and the Lab->RGB coefficients are built inside.
 * Replaces: the implicit table initialisation inside cv2.cvtColor, reached from
 * mdir/components/data/transform/functional.py:35,63.
 */
int gdt_init(const int16_t* host_rgb2lab_lut);
int gdt_is_initialised(void);
/* debug/test hook: copy the 1024x4 inverse-gamma spline table built by gdt_init to host memory */
int gdt_debug_get_spline_table(float* host_out_4096);

/* debug/test hook: K1 replaces `(x - mean) / std` by a divider-free, correctly rounded sequence; this counts (into the
 * device counter, which the caller zeroes) the floats a with bit pattern in [lo_bits, hi_bits], both signs, for which
 * that sequence differs from IEEE a / b. Expected: 0. */
int gdt_debug_div_check(float b, uint32_t lo_bits, uint32_t hi_bits, unsigned long long* mismatches_dev, void* stream);

/* debug/test hook: work split and pipe choice of K1's table lookups (every combination computes bit-identical results).
 *   texab    : bit 0: pass A fetches the chroma lattice records through the texture pipe (when chroma_a);
 *              bit 1 / bit 2: pass A fetches all / every other lightness record through the texture pipe (when !chroma_a)
 *   spltex   : 0..1 of pass B's inverse-gamma spline lookups go through the texture pipe
 *   fytex    : pass B takes the lightness half of Lab->RGB from a 256-entry table through the texture pipe
 *   chroma_a : the chroma is interpolated in pass A (one lattice visit per pixel) instead of pass B; -1 = automatic
 *              (pass A, from the compressed record, whenever that is available: the default)
 *   occ_a    : resident CTAs per SM pass A is compiled for (4 or 6) */
int gdt_debug_k1_config(int texab, int spltex, int fytex, int chroma_a, int occ_a);
/* debug/test hook: force the image rows per pass-B CTA of K1 (1..64; 0 = the built-in wave-quantisation rule) */
int gdt_debug_k1_rows(int rows_per_cta);
/* debug/test hook: pass B of K1 with packed f32x2 arithmetic (two pixels per instruction) or scalar (0, the default);
 * bit-identical results */
int gdt_debug_k1_pack(int packed_f32x2);
/* debug/test hook: pass B of K1 as one persistent 1024-thread CTA per SM with eight conflict-free shared-memory copies of
 * the inverse-gamma spline and of the lightness table (1, the default: on widths without OpenCV scalar-tail pixels;
 * 2: on every width) or as one 256-thread CTA per row band (0); bit-identical results */
int gdt_debug_k1_persist(int persistent);
/* debug/test hook: the persistent pass B normalises with ONE correction step of the divider-free division for the std
 * values verified exhaustively at gdt_init (default 1) or always with two (0); bit-identical results.
 * gdt_debug_k1_div1_verified: 1 when `std` passed that check on the current device, 0 when not, < 0 on error */
int gdt_debug_k1_div1(int one_step);
/* debug/test hook: pass A hands pass B the chroma as the two float terms of Lab->RGB (8 B/px scratch; 1, on the common
 * sizes only) or as the Q14 pair (0, the default: measured faster); bit-identical results */
int gdt_debug_k1_chroma_f(int float_terms);
int gdt_debug_k1_div1_verified(float std);
/* debug/test hook: pass A interpolates from the compressed 32-byte lattice record (one sector gather per pixel; default 1
 * when the table fits the format) or from the uncompressed 16 + 32-byte records (0); bit-identical results */
int gdt_debug_k1_rec32(int compressed_record);
/* debug/test hook: images per (pass A, pass B) launch pair of K1: 0 = the whole batch at once (default), -1 = sized so
 * that a chunk's 5 B/px scratch stays L2-resident between the passes, > 0 = forced */
int gdt_debug_k1_chunk(int images_per_launch_pair);

/* ---- K5: dataset image geometry (crop + LANCZOS thumbnail) ---------------------------------------
 * The image-size half of the reference's dataset loader on the device:
 *   ImagesFromList.__getitem__ crop + imresize   (mdir/external/cirtorch/datasets/genericdataset.py:86-97)
 *   imresize = img.thumbnail((imsize, imsize), Image.LANCZOS)   (mdir/external/cirtorch/datasets/datahelpers.py:75-82)
 * Bit-exact against Pillow's 8-bit resampler (aspect-preserving target size, reducing_gap = 2.0 integer box reduction,
 * two-pass 22-bit fixed-point LANCZOS with uint8 rounding between the passes). A crop is a sub-rectangle of the source:
 * pass the pointer of its first pixel and the full image's row stride.
 *
 * gdt_thumbnail_geometry : host only. Target size and pre-reduction factors Pillow picks for a w x h image and the request
 *                          (imsize, imsize); returns 1 when the image is resized, 0 when it is left alone (out = in),
 *                          negative gdt_status on bad arguments.
 * gdt_resize_plan_create : per (w, h, imsize) geometry: filter coefficients computed on the host in double precision
 *                          (the libm calls Pillow makes) and uploaded to the current device. Synchronous; setup, not
 *                          hot path. Destroy with gdt_resize_plan_destroy.
 * gdt_resize_u8          : src uint8 [h][w][3] with row stride src_stride bytes (device) -> dst uint8
 *                          [out_h][out_w][3] contiguous (device). Asynchronous on `stream`; scratch from `ws`
 *                          (gdt_resize_workspace_bytes(plan), 256-byte aligned). 1 - 3 launches. */
typedef struct gdt_resize_plan gdt_resize_plan;
int gdt_thumbnail_geometry(int w, int h, double imsize, int* out_w, int* out_h, int* fx, int* fy);
int gdt_resize_plan_create(int in_w, int in_h, double imsize, gdt_resize_plan** plan_out);
void gdt_resize_plan_destroy(gdt_resize_plan* plan);
int gdt_resize_plan_info(const gdt_resize_plan* plan, int* out_w, int* out_h, int* fx, int* fy);
size_t gdt_resize_workspace_bytes(const gdt_resize_plan* plan);
int gdt_resize_u8(const gdt_resize_plan* plan, const uint8_t* src, size_t src_stride, uint8_t* dst, void* ws,
                  size_t ws_bytes, void* stream);
/* Batched form: n decoded images of ONE geometry (one plan), each with its own base pointer and row stride (host arrays
 * of device pointers / strides), resized by one launch per pass; dst = uint8 [n][out_h][out_w][3] contiguous. */
size_t gdt_resize_batch_workspace_bytes(const gdt_resize_plan* plan, int n);
int gdt_resize_u8_batch(const gdt_resize_plan* plan, const uint8_t* const* host_srcs, const size_t* host_strides, int n,
                        uint8_t* dst, void* ws, size_t ws_bytes, void* stream);
/* N1, decode stage (opt-in): batches of JPEG bit streams decoded on the GPU by the nvJPEG LIBRARY (hardware JPEG engines
 * when the device exposes them, nvJPEG's default backend otherwise); replaces `pil_loader`
 * (mdir/external/cirtorch/datasets/datahelpers.py:20-27) for JPEG files. Library code, NOT bit-identical to libjpeg (a few
 * grey levels): the default loader keeps PIL decoding. libnvjpeg is opened lazily (dlopen); without it these return
 * GDT_ERR_UNSUPPORTED and everything else works.
 *   gdt_jpeg_available     1 when libnvjpeg could be loaded
 *   gdt_jpeg_dims          host-side header parse -> width, height of the decoded image
 *   gdt_jpeg_decode_batch  jpegs[i] / nbytes[i]: HOST bit streams; dev_rgb[i]: DEVICE buffer of heights[i] x widths[i] x 3
 *                          bytes (interleaved RGB, what gdt_resize_u8 / gdt_clahe_u8 consume); enqueued on `stream`
 *   gdt_debug_jpeg_last_backend  backend of the last batch: 1 = hardware engines, 2 = GPU-hybrid (batches of >= 50 baseline
 *                          streams: Huffman decoding on the GPU), 3 = default (Huffman decoding on the host's threads)
 *   gdt_debug_jpeg_status  nvjpegStatus_t of {hardware create, hardware batch init, hardware decode, fallback decode} */
int gdt_jpeg_available(void);
int gdt_jpeg_dims(const uint8_t* jpeg, size_t nbytes, int* width, int* height);
int gdt_jpeg_decode_batch(const uint8_t* const* jpegs, const size_t* nbytes, int n, uint8_t* const* dev_rgb,
                          const int* widths, const int* heights, void* stream);
int gdt_debug_jpeg_last_backend(void);
int gdt_debug_jpeg_status(int* out4);
/* debug/test hook: 1 = K5 runs its byte-wise kernels only (the dp4a kernels off), for A/B timing and parity */
int gdt_debug_k5_bytewise(int on);
/* debug/test hook: 0 = K5's horizontal pass runs the interleaved dp4a kernel instead of the planar one (A/B, parity) */
int gdt_debug_k5_planar(int on);
/* debug/test hook (host only): Pillow's precompute_coeffs + normalize_coeffs_8bpc for the LANCZOS filter.
 * bounds: out_size x (first input index, tap count); kk: out_size x ksize int32 (capacity in elements). */
int gdt_debug_resize_coeffs(int in_size, float in0, float in1, int out_size, int* ksize, int* bounds, int32_t* kk,
                            size_t kk_capacity);

/* ---- K1: CLAHE preprocessing ------------------------------------------------------------------
 * Fused `pil2np | apply_clahe:clip:grid:lab | totensor | normalize`
 * (mdir/components/data/transform/core_transforms.py:35-100,
 *  mdir/components/data/transform/photometric_transforms.py:28-36,
 *  mdir/components/data/transform/functional.py:28-35,55-63,81-85,140-161).
 * rgb_hwc : n images, uint8, [n][h][w][3]          out_chw : float32 [n][3][h][w]
 * host_mean/host_std : 3 floats each (Normalize). Bit-exact against the reference's OpenCV path.
 */
size_t gdt_clahe_workspace_bytes(int n, int h, int w, int grid);
int gdt_clahe_u8(const uint8_t* rgb_hwc, int n, int h, int w, double clip_limit, int grid,
                 const float* host_mean, const float* host_std, float* out_chw,
                 void* ws, size_t ws_bytes, void* stream);

/* `ClahePost.postprocess` (mdir/components/data/wrapper.py:325-348): CLAHE on a normalised float
 * CHW device tensor without the device->host->device round trip.
 * in_chw : float32 [n][3][h][w] normalised with (in_mean, in_std); result re-normalised with
 * (out_mean, out_std) (the reference uses the same pair for both). */
int gdt_clahe_f32(const float* in_chw, int n, int h, int w, double clip_limit, int grid,
                  const float* host_in_mean, const float* host_in_std,
                  const float* host_out_mean, const float* host_out_std, float* out_chw,
                  void* ws, size_t ws_bytes, void* stream);

/* `MeanStdPost._adapt` / `MeanStdPre` (mdir/components/data/wrapper.py:149-194): re-normalisation of a CHW float tensor,
 *     y = ((x * in_std[c] + in_mean[c]) - out_mean[c]) / out_std[c]
 * with the reference's four separately rounded fp32 operations (mul, add, sub, true division), bit-exact against the
 * torch expression `x.mul(s0).add(m0).sub(m1).div(s1)`. x, y : float32 [n][3][plane] (plane = h*w), may alias. */
int gdt_meanstd_adapt(const float* x, long long n, long long plane, const float* host_in_mean, const float* host_in_std,
                      const float* host_out_mean, const float* host_out_std, float* y, void* stream);

/* ---- K2: GeM + L2N + multi-scale aggregation + learned whitening ---------------------------------
 * Replaces LF.gem / LF.l2n (mdir/external/cirtorch/layers/functional.py:21-22,130-131),
 * CirMultiscaleAggregation.aggregate_tensor/postprocess (mdir/components/data/wrapper.py:235-260)
 * and CirtorchWhiten.postprocess (mdir/components/data/wrapper.py:320-322).
 *
 * host_fmaps[s] : device pointer to the backbone's final feature map of scale s,
 *                 float32 [n][c][host_h[s]][host_w[s]] contiguous, s < scales (<= GDT_MAX_SCALES)
 * p_dev         : device pointer to GeM's 1-element exponent `pool.p` (read on the device: no sync)
 * eps           : GeM clamp (1e-6)
 * flags         : GDT_GEM_AGGREGATE   apply the multi-scale generalised mean + eps-free renorm
 *                                    (cirmultiscale wrapper present; also valid for scales == 1)
 *                 GDT_GEM_MSP_IS_P    exponent of that mean is p (wrapper.py:248-251), else 1
 * P, m          : whitening projection [dim][c] row-major (row stride ldP) and mean [c], or NULL
 * P_split       : optional [2][dim][c] output of gdt_whiten_prepare(P) (TF32 hi / lo halves of the learned, constant
 *                 projection, prepared once per whitening): selects the tcgen05 (kind::tf32, 3xTF32) projection
 *                 kernel; NULL selects the mma.sync 3xTF32 kernel. Both are fp32-accurate (~1e-6 relative).
 * desc          : float32 [n][dim] row-major (dim == c when P == NULL)
 */
#define GDT_MAX_SCALES 8
#define GDT_GEM_AGGREGATE 1
#define GDT_GEM_MSP_IS_P 2
#define GDT_DESC_NORMALISED 4 /* internal to gdt_desc_post: inputs are already L2-normalised descriptors */
#define GDT_POOLED_RAW_MEAN 8 /* internal to gdt_gem_whiten: pooled values are means, the 1/p root is still due */
#define GDT_SPLIT_OUT 16      /* internal: centred descriptors are written as TF32 hi / lo halves */
size_t gdt_gem_whiten_workspace_bytes(int n, int c, int scales, int dim);
int gdt_gem_whiten(const float* const* host_fmaps, const int* host_h, const int* host_w,
                   int n, int c, int scales, const float* p_dev, float eps, int flags,
                   const float* P, int ldP, const float* P_split, const float* m, int dim,
                   float* desc, void* ws, size_t ws_bytes, void* stream);
int gdt_whiten_prepare(const float* P, int ldP, int c, int dim, float* P_split, void* stream);

/* The module-level pieces of the same path, for callers that drive the reference's objects one by one:
 *   gdt_gem_pool  `GeM.forward` / LF.gem (layers/pooling.py:36-47, layers/functional.py:21-22):
 *                 fmap [n][c][h][w] -> pooled [n][c] (no normalisation)
 *   gdt_l2n_rows  `L2N.forward` / LF.l2n (layers/normalization.py:10-20, layers/functional.py:130-131):
 *                 out[r] = x[r] / (||x[r]||_2 + eps), x/out [n][dim] (in place allowed)
 *   gdt_desc_post `CirMultiscaleAggregation.aggregate_tensor` (wrapper.py:235-245) and/or
 *                 `CirtorchWhiten.postprocess` (wrapper.py:320-322) on already L2-normalised descriptors:
 *                 host_descs[s] : device pointers to [n][c] per-scale descriptors; exponent of the
 *                 generalised mean = msp_dev[0] when flags has GDT_GEM_MSP_IS_P (device read, no sync), else
 *                 msp_host; GDT_GEM_AGGREGATE as in gdt_gem_whiten; P/m/dim as in gdt_gem_whiten. */
int gdt_gem_pool(const float* fmap, int n, int c, int h, int w, const float* p_dev, float eps, float* pooled,
                 void* stream);
int gdt_l2n_rows(const float* x, int n, int dim, float eps, float* out, void* stream);
size_t gdt_desc_post_workspace_bytes(int n, int c, int dim);
int gdt_desc_post(const float* const* host_descs, int n, int c, int scales, const float* msp_dev, float msp_host,
                  int flags, const float* P, int ldP, const float* P_split, const float* m, int dim, float* out,
                  void* ws, size_t ws_bytes, void* stream);

/* ---- K3: query x database scoring fused with streaming top-k ----------------------------------------
 * Replaces `scores = np.dot(vecs.T, qvecs); ranks = np.argsort(-scores, axis=0)` restricted to the
 * first k ranks (mdir/components/optim/score/cirscore.py:71-72).
 *
 * Database shards are prepared once (gdt_db_prepare): an fp16 shadow copy (rows scaled by an exact power of
 * two) feeds the tcgen05 coarse pass, the fp32 rows feed the exact re-scoring of the few survivors, so the
 * returned scores are fp32-exact (fp64-accumulated dot, rounded once) and the ranking is exact:
 * ordering = (score descending, global index ascending). The coarse pass only filters, with a margin derived
 * from the measured fp16 rounding residuals (see score_topk_sm100.cu).
 *
 * q        : float32 [nq][d] row-major queries            db : float32 [ndb][d] row-major shard
 * db_f16   : shadow written by gdt_db_prepare, [ndb][d] fp16 (16-byte aligned)
 * db_stats : device float[4] written by gdt_db_prepare: max row norm, power-of-two scale, max norm of the
 *            rounding residual, max norm of the shadow rows
 * top_scores/top_idx : [nq][k]; top_idx holds index_base + local row; when ndb < k the tail is
 *            filled with (-inf, -1)
 * status_dev : device int32[4], zero on success; [0] = GDT_ERR_CANDIDATE_OVERFLOW if a query's
 *            candidate set did not fit (that query's top_idx[0] is -2 and its results must be recomputed with
 *            gdt_score_topk_exact), [1] = largest number of candidates re-scored for one query,
 *            [2] = number of overflowed queries, [3] = largest number of coarse candidates of one query
 * Requirements: d % 8 == 0, d <= 8192, k <= 1024, index_base + ndb <= 2^32.
 */
size_t gdt_db_prepare_workspace_bytes(long long ndb, int d);
int gdt_db_prepare(const float* db, long long ndb, int d, void* db_f16, float* db_stats,
                   void* ws, size_t ws_bytes, void* stream);
/* Two-phase form for row-sharded databases (the one real exchange step of the path). Every rank must
 * filter in the same score units, so the shard statistics are made common first:
 *     gdt_db_prepare_norm -> all-reduce(MAX) db_stats[0] -> gdt_db_prepare_convert -> all-reduce(MAX) db_stats[2..3]
 * and per search
 *     gdt_score_topk_filter -> all-reduce(SUM) of the per-query score histograms -> gdt_score_topk_finalize.
 * With the summed histogram each rank keeps only candidates above the GLOBAL k-th-score threshold, so a shard
 * re-scores ~k/G instead of ~k survivors per query and may return fewer than k valid entries (padding (-inf, -1));
 * gdt_topk_merge over the gathered lists then yields exactly the global top k.
 * gdt_score_topk_exchange_layout reports where the uint32 [nq][256] histograms live inside the workspace. */
int gdt_db_prepare_norm(const float* db, long long ndb, int d, float* db_stats, void* stream);
int gdt_db_prepare_convert(const float* db, long long ndb, int d, void* db_f16, float* db_stats, void* stream);
int gdt_score_topk_exchange_layout(int nq, long long ndb, int d, int k, size_t* hist_offset, size_t* hist_bytes);
int gdt_score_topk_filter(const float* q, const void* db_f16, const float* db_stats, int nq, long long ndb, int d, int k,
                          int32_t* status_dev, void* ws, size_t ws_bytes, void* stream);
int gdt_score_topk_finalize(const float* q, const float* db, int nq, long long ndb, int d, int k, long long index_base,
                            float* top_scores, int64_t* top_idx, int32_t* status_dev,
                            void* ws, size_t ws_bytes, void* stream);
/* debug/test hook: gdt_score_topk_filter that also writes every raw tensor-core (coarse) score to coarse[nq][ndb] and the
 * per-query filter constants to meta_out[nq][4] = {scale, 1/scale, margin = 2*E_q, sq}. The exactness of K3 rests on
 * |coarse - sq*sx*<q,x>| <= E_q; tests/test_gpu_topk.py measures the left side against fp64 on adversarial data. */
int gdt_debug_k3_coarse_scores(const float* q, const void* db_f16, const float* db_stats, int nq, long long ndb, int d,
                               int k, float* coarse, float* meta_out, int32_t* status_dev, void* ws, size_t ws_bytes,
                               void* stream);
size_t gdt_score_topk_workspace_bytes(int nq, long long ndb, int d, int k);
int gdt_score_topk(const float* q, const float* db, const void* db_f16, const float* db_stats,
                   int nq, long long ndb, int d, int k, long long index_base,
                   float* top_scores, int64_t* top_idx, int32_t* status_dev,
                   void* ws, size_t ws_bytes, void* stream);

/* Exact CUDA-core variant of the same contract (no fp16 shadow, no tensor cores); used for small
 * problems and as an on-device cross-check of the tcgen05 path. */
size_t gdt_score_topk_exact_workspace_bytes(int nq, long long ndb, int d, int k);
int gdt_score_topk_exact(const float* q, const float* db, int nq, long long ndb, int d, int k,
                         long long index_base, float* top_scores, int64_t* top_idx,
                         void* ws, size_t ws_bytes, void* stream);

/* Merge g per-shard top-k lists (after the NCCL allgather) into one:
 * scores/idx : [g][nq][k] -> out : [nq][k], same ordering rule; entries with idx < 0 are padding. */
int gdt_topk_merge(const float* scores, const int64_t* idx, int g, int nq, int k,
                   float* out_scores, int64_t* out_idx, void* stream);

/* Packed exchange form of the per-shard lists: one uint64 per entry = the library's rank key of (score, GLOBAL index)
 * (larger key sorts first; 0 = padding, 2 = "this shard overflowed"), so that the row-sharded search needs ONE
 * all-gather of 8 bytes per entry instead of two (fp32 scores + int64 indices, 12 bytes).
 *   gdt_topk_pack          scores/idx [n] -> keys [n]; idx -1 -> padding, idx -2 (overflow marker) -> overflow key
 *   gdt_topk_merge_packed  keys [g][nq][k] -> merged (scores, idx) [nq][k]; a query for which any shard reported an
 *                          overflow comes back with idx[q][0] == -2, as gdt_score_topk_finalize marks it. */
int gdt_topk_pack(const float* scores, const int64_t* idx, long long n, uint64_t* keys, void* stream);
int gdt_topk_merge_packed(const uint64_t* keys, int g, int nq, int k, float* out_scores, int64_t* out_idx, void* stream);
/* Query-sharded merge (the row-sharded search on many GPUs): every rank merges only its slice of the queries -- the
 * per-shard lists reach it by an all-to-all instead of an all-gather (1 / world of the bytes and of the merge work) --
 * and the merged slices are gathered in packed form:
 *   gdt_topk_merge_packed_keys  keys [g][nq][k] -> merged KEYS [nq][k] (0 = padding; out[q][0] == 2 marks an overflow)
 *   gdt_topk_unpack             keys [n] -> (scores, idx) [n]: padding -> (-inf, -1), the overflow key -> (-inf, -2) */
int gdt_topk_merge_packed_keys(const uint64_t* keys, int g, int nq, int k, uint64_t* out_keys, void* stream);
int gdt_topk_unpack(const uint64_t* keys, long long n, float* scores, int64_t* idx, void* stream);

/* ---- K4: ranks of ground-truth ids + mAP ------------------------------------------------------------
 * Replaces the use of the full `ranks` matrix inside compute_map
 * (mdir/external/cirtorch/utils/evaluate.py:39-111): for every query and every probe id (its
 * positives and junk), the 0-based position the id would have in the full descending ranking
 * = number of database rows that sort before it.
 *
 * probe_idx  : int64 [nq][pmax] global ids (index_base-relative rows live on this shard), -1 = pad
 * probe_score: float32 [nq][pmax] exact score of each probe (gdt_probe_scores on the owning shard,
 *              summed across shards by the caller since non-owners contribute 0)
 * before     : int64 [nq][pmax], += number of rows of THIS shard that sort before the probe
 *              (caller zero-initialises and sums across shards)
 * The probes of a query are sorted once; every (row, query) pair then costs one binary search and one counter
 * increment, whatever pmax is (pmax <= 2048, global ids < 2^32). Probes with equal (score, id) get equal counts.
 */
int gdt_probe_scores(const float* q, const float* db, int nq, long long ndb, int d,
                     long long index_base, const int64_t* probe_idx, int pmax,
                     float* probe_score, void* stream);
size_t gdt_rank_counts_workspace_bytes(int nq, int pmax);
/* debug/test hook: 1 = gdt_rank_counts re-scores every (row, query) pair exactly (fp64-accumulated) instead of ranking from
 * the fp32 score + error bound; the counts must not change */
int gdt_debug_k4_exact(int on);
int gdt_rank_counts(const float* q, const float* db, int nq, long long ndb, int d,
                    long long index_base, const int64_t* probe_idx, const float* probe_score,
                    int pmax, int64_t* before, void* ws, size_t ws_bytes, void* stream);

/* compute_ap / compute_map body (evaluate.py:3-37,60-106) for nq queries:
 * pos_rank [nq][pmax_pos], junk_rank [nq][pmax_junk] : 0-based full-ranking positions (any order,
 * first npos[q] / njunk[q] entries valid). nres [nq] (nullable -> npos): length of the positive list as given, the
 * recall denominator `compute_ap(pos, len(qgnd))` uses (evaluate.py:75-98; differs from npos when the list holds
 * duplicated or foreign ids, which np.in1d finds once or never). kappas [nk] (device). Outputs (float64, as the
 * reference): ap [nq] (NaN when nres == 0 or npos == 0), prk [nq][nk]. Several protocols (easy / medium / hard) are
 * evaluated in one launch by stacking their rows. */
int gdt_map_eval(const int64_t* pos_rank, int pmax_pos, const int64_t* junk_rank, int pmax_junk,
                 const int32_t* npos, const int32_t* njunk, const int32_t* nres, int nq,
                 const int32_t* kappas, int nk, double* ap, double* prk, void* stream);

/* ---- N3: diverse-anchor mining -------------------------------------------------------
 * The greedy loop of `DiverseAnchorsDataset._select_positive_pairs_db`
 * (mdir/components/data/dataset/cirtorch_datasets.py:68-100) without a host round trip per round:
 *   pool      float32 [n][d] rows (the reference's qvecs is the D x n transpose)
 *   ranks_dev int32 [steps]: the ascending rank (under value asc, index asc) of `most_similar` to pick in each round --
 *             `dissimilar_split + choice` of the reference, data-independent, computed by the caller
 *   first     index picked before the first round (the reference starts from 0)
 * Outputs: picked_dev int32 [steps + 1] (first, then one index per round), picked_score_dev float32 [steps] (the
 * reference's qscore_acc). Scores are the library's exact fp32 scores (fp64-accumulated). ws: n floats. */
size_t gdt_diverse_anchors_workspace_bytes(int n);
int gdt_diverse_anchors(const float* pool, int n, int d, const int32_t* ranks_dev, int steps, int first,
                        int32_t* picked_dev, float* picked_score_dev, void* ws, size_t ws_bytes, void* stream);

/* ---- N2: whitening learning ------------------------------------------------------------
 * The symmetric rank-n updates of `whitenlearn` / `pcawhitenlearn` (mdir/external/cirtorch/utils/whiten.py:14-50:
 * `np.dot(df, df.T)`, `np.dot(Xc, Xc.T)`) in float64:  C [d][d] = alpha * A A^T  for row-major A [d][n] with row stride
 * lda. Lower-triangular tiles only, mirrored; the split of n is summed in a fixed order (bit-reproducible). */
size_t gdt_syrk_f64_workspace_bytes(int d, long long n);
int gdt_syrk_f64(const double* A, int d, long long n, long long lda, double alpha, double* C, void* ws, size_t ws_bytes,
                 void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GANDTR_B200_H */
