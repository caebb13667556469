/*
 * gsr.h — C ABI of the B200-native differentiable Gaussian-splatting rasterizer (libgsr_b200.so).
 *
 * This is the drop-in boundary for GS-LIVM's `src/cuda_rasterizer` hot path. Every entry point takes
 * plain pointers and sizes only (no torch / C++ types) and replaces one reference interface:
 *
 *   gsr_forward       <- CudaRasterizer::Rasterizer::forward    (include/gs/cuda_rasterizer/rasterizer.h:23-51,
 *                                                                src/cuda_rasterizer/rasterizer_impl.cu:181-342)
 *   gsr_backward      <- CudaRasterizer::Rasterizer::backward   (rasterizer.h:53-88, rasterizer_impl.cu:346-457)
 *   gsr_mark_visible  <- CudaRasterizer::Rasterizer::markVisible(rasterizer.h:21, rasterizer_impl.cu:128-135)
 *   gsr_required_*    <- CudaRasterizer::required<GeometryState|ImageState|BinningState>
 *                                                               (include/gs/cuda_rasterizer/rasterizer_impl.h:61-66)
 *   gsr_view_*        <- GeometryState/ImageState/BinningState::fromChunk (rasterizer_impl.cu:137-177);
 *                        debug/parity accessors only — the layout inside the three byte buffers is OURS.
 *
 * The C++ class `CudaRasterizer::Rasterizer` with the reference's exact signatures is provided on top of
 * this ABI by include/cuda_rasterizer/rasterizer.h + csrc/rasterizer_shim.cpp (same library), so the
 * reference's src/gs/rasterize_points.cu and src/gs/rasterizer.cu compile and link unchanged.
 *
 * Conventions
 *  - All data pointers are DEVICE pointers on the current CUDA device, FP32 / INT32, contiguous, laid out
 *    exactly as the reference expects them (means3D[P,3], shs[P,M,3], out_color[3,H,W] planar, ...).
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream, what the reference uses).
 *  - Functions return GSR_OK (0) or a negative gsr_status; they never throw. gsr_last_error() returns a
 *    thread-local human-readable message for the last failure on the calling thread.
 *  - The three allocator callbacks are invoked once each, in the order geometry -> image -> binning
 *    (binning only after the number of rendered instances R is known), like the reference does
 *    (rasterizer_impl.cu:213-215,225-227,279-281). The memory must stay valid until the matching backward.
 *  - There is NO CPU fallback: without a CUDA device every compute entry point returns GSR_ERR_CUDA.
 */
#ifndef GSR_B200_H_INCLUDED
#define GSR_B200_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define GSR_API __attribute__((visibility("default")))
#else
#define GSR_API
#endif

#define GSR_ABI_VERSION 5
#define GSR_NUM_CHANNELS 3 /* NUM_CHANNELS, include/gs/cuda_rasterizer/config.h:15 */
#define GSR_BLOCK_X 16     /* BLOCK_X, config.h:16 */
#define GSR_BLOCK_Y 16     /* BLOCK_Y, config.h:17 */

typedef enum gsr_status {
  GSR_OK = 0,
  GSR_ERR_INVALID = -1, /* bad argument combination (cf. std::runtime_error, rasterizer_impl.cu:229-231) */
  GSR_ERR_CUDA = -2,    /* a CUDA runtime call or kernel failed */
  GSR_ERR_ALLOC = -3    /* an allocator callback returned NULL */
} gsr_status;

/* Scratch allocator: must return `bytes` bytes of device memory (any alignment; we align to 128 inside). */
typedef char* (*gsr_alloc_fn)(void* ctx, size_t bytes);

/* Arguments of one forward render; field meanings == reference Rasterizer::forward parameters. */
typedef struct gsr_fwd_args {
  int P;                       /* number of Gaussians */
  int D;                       /* active SH degree (0..3) */
  int M;                       /* SH coefficients per Gaussian per channel stored in `shs` (0 if unused) */
  int width, height;           /* image size in pixels */
  const float* background;     /* [3] */
  const float* means3D;        /* [P,3] */
  const float* shs;            /* [P,M,3] or NULL when colors_precomp is given */
  const float* colors_precomp; /* [P,3] or NULL -> colours from SH */
  const float* opacities;      /* [P] activated */
  const float* scales;         /* [P,3] activated, or NULL when cov3D_precomp is given */
  float scale_modifier;
  const float* rotations;      /* [P,4] (r,x,y,z); NOT normalised inside (reference forward.cu:146) */
  const float* cov3D_precomp;  /* [P,6] or NULL */
  const float* viewmatrix;     /* [16] column-major as the reference indexes it (auxiliary.h:48-55) */
  const float* projmatrix;     /* [16] */
  const float* cam_pos;        /* [3] */
  float tan_fovx, tan_fovy;
  int prefiltered;             /* accepted for signature parity; unused (reference preprocess ignores it) */
  float* out_color;            /* [3,H,W] */
  float* out_depth;            /* [1,H,W]  un-normalised sum z*alpha*T (forward.cu:386) */
  float* out_acc;              /* [1,H,W]  silhouette sum alpha*T      (forward.cu:387) */
  int* radii;                  /* [P] or NULL -> internal */
  int debug;                   /* !=0: synchronise + check after every stage (CHECK_CUDA, auxiliary.h:146-154) */
  int activations;             /* 0: opacities / scales / rotations are activated values, as the reference passes them.
                                * GSR_ACT_* bits: they are the RAW optimised parameters and the kernels apply the
                                * reference's activations themselves */
} gsr_fwd_args;

/* gsr_fwd_args.activations / gsr_bwd_args.activations: GaussianModel's activations (include/gs/gs/gaussian.cuh:40-54)
 * fused into the kernels instead of running as LibTorch elementwise ops (and their autograd backwards) around every
 * render. With a bit set the corresponding input is the raw parameter, the forward applies
 *   opacity = sigmoid(_opacity) = 1 / (1 + exp(-x)),  scale = exp(_scaling),  rotation = _rotation / max(|_rotation|, 1e-12)
 * on load, and the backward returns the gradient w.r.t. the RAW parameter in the same output array
 * (dL_dopacity, dL_dscale, dL_drot). The SH coefficients need no entry point of their own: pass ONE [P,M,3] array
 * (what torch::cat({_features_dc, _features_rest}, 1) would build, render_utils.cuh:38-44) that the optimiser updates
 * in place -- gsr_adam_step takes the two learning rates as a pattern over the row (gsr_adam_group.period). */
#define GSR_ACT_OPACITY_SIGMOID 1
#define GSR_ACT_SCALE_EXP 2
#define GSR_ACT_ROTATION_NORMALIZE 4

/* Arguments of one backward; == reference Rasterizer::backward parameters. Gradient outputs must be
 * pre-zeroed by the caller (rasterize_points.cu:173-181); there is no dL_ddepth input (rasterizer.cu:77-80). */
typedef struct gsr_bwd_args {
  int P, D, M, R;
  int width, height;
  const float* background;
  const float* means3D;
  const float* shs;
  const float* colors_precomp;
  const float* scales;
  float scale_modifier;
  const float* rotations;
  const float* cov3D_precomp;
  const float* viewmatrix;
  const float* projmatrix;
  const float* cam_pos;
  float tan_fovx, tan_fovy;
  const int* radii;      /* [P] as written by the forward, or NULL -> internal */
  char* geom_buffer;     /* the three buffers the forward's allocators returned */
  char* binning_buffer;
  char* image_buffer;
  const float* dL_dpix;  /* [3,H,W] */
  const float* dL_dacc;  /* [1,H,W] */
  float* dL_dmean2D;     /* [P,3] (z unused) */
  float* dL_dconic;      /* [P,2,2] (index 2 unused) */
  float* dL_dopacity;    /* [P,1] */
  float* dL_dcolor;      /* [P,3] */
  float* dL_dmean3D;     /* [P,3] */
  float* dL_dcov3D;      /* [P,6] */
  float* dL_dsh;         /* [P,M,3] */
  float* dL_dscale;      /* [P,3] */
  float* dL_drot;        /* [P,4] */
  int debug;
  int flags;             /* 0 = the reference's contract (all 9 outputs pre-zeroed by the caller); GSR_BWD_* bits */
  unsigned char* row_state; /* optional, with GSR_BWD_OUTPUTS_UNINITIALIZED: P bytes owned by the caller, all 1 when the
                            * five assigned outputs hold arbitrary data (e.g. freshly allocated), kept by the library
                            * from then on: != 0 <=> the outputs of that Gaussian may be non-zero. A culled Gaussian
                            * whose outputs are still zero from an earlier backward is then skipped instead of being
                            * zero-filled again (256 B per Gaussian per frame). NULL: every culled Gaussian is written.
                            * The same five arrays must be passed every time; with GSR_BWD_SH_FACTORED the state covers
                            * the other four. */
  int activations;       /* GSR_ACT_* bits as given to the forward; scales / rotations / opacities are then the raw
                          * parameters and dL_dscale / dL_drot / dL_dopacity are gradients w.r.t. them */
  const float* opacities; /* [P]; only read with GSR_ACT_OPACITY_SIGMOID */
  int range_first;       /* with GSR_BWD_ONLY_GEOM: the per-Gaussian kernel covers Gaussians [range_first, range_first + */
  int range_count;       /* range_count) only (range_count == 0: all P) -- a window exchange pipelines it in chunks */
} gsr_bwd_args;

/* gsr_bwd_args.flags. With GSR_BWD_OUTPUTS_UNINITIALIZED the caller does NOT zero the gradient outputs (the
 * reference's caller does, 300 B per Gaussian per backward, rasterize_points.cu:173-181): the library writes
 * every element of the five assigned outputs (dL_dmean3D, dL_dcov3D, dL_dsh, dL_dscale, dL_drot; zeros for
 * Gaussians with radii <= 0) and zero-fills the four accumulated ones (dL_dmean2D, dL_dconic, dL_dopacity,
 * dL_dcolor) itself on `stream`. Same results, 256 B per Gaussian less memset traffic. */
#define GSR_BWD_OUTPUTS_UNINITIALIZED 1
/* with the flag above: the caller already ran gsr_backward_prepare() on these outputs (e.g. on another stream, beside
 * the forward), so gsr_backward() must not clear the accumulated arrays again */
#define GSR_BWD_ACCUMULATORS_ZEROED 2
/* dL_dsh is NOT written (the pointer may be NULL): per view, dL_dsh[k][c] = basis_k(dir) * dRGB[c] has rank one, and a
 * view-sharded window exchanges the three dRGB floats instead of the 48 of the row and rebuilds the sum of the rows on
 * every rank (gsr_window_push_rgb / gsr_window_expand_sh below). Everything else is computed as usual. */
#define GSR_BWD_SH_FACTORED 4
/* one half of the backward per call, so that a caller can start work that needs only the blend backward's outputs
 * (dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor are complete after it) beside the per-Gaussian half:
 * GSR_BWD_ONLY_BLEND = clearing of the accumulators + blend backward; GSR_BWD_ONLY_GEOM = the per-Gaussian kernel. */
#define GSR_BWD_ONLY_BLEND 8
#define GSR_BWD_ONLY_GEOM 16

/* Debug / parity view of the opaque buffers. Strides are in BYTES; count = number of elements. */
typedef struct gsr_array_view {
  void* ptr;
  size_t stride;
  size_t count;
} gsr_array_view;

typedef struct gsr_geom_view {
  gsr_array_view means2D;       /* float2 per Gaussian */
  gsr_array_view extent;        /* float2 per Gaussian: conservative |dx|,|dy| bound of alpha>=1/255 (ours) */
  gsr_array_view conic_opacity; /* float4 per Gaussian */
  gsr_array_view rgb;           /* float3 per Gaussian */
  gsr_array_view depths;        /* float per Gaussian */
  gsr_array_view clamped;       /* uint8 bitmask per Gaussian: bit c = channel c clamped */
  gsr_array_view internal_radii;/* int32 per Gaussian */
  gsr_array_view tiles_touched; /* uint32 per Gaussian */
  gsr_array_view point_offsets; /* uint32 per Gaussian (inclusive scan) */
} gsr_geom_view;

typedef struct gsr_binning_view {
  gsr_array_view point_list_keys_unsorted; /* uint64 */
  gsr_array_view point_list_unsorted;      /* uint32 */
  gsr_array_view point_list_keys;          /* uint64, sorted */
  gsr_array_view point_list;               /* uint32, sorted */
  gsr_array_view instance_records;         /* 48-byte packed per-instance records in sorted order (ours) */
  gsr_array_view block_candidates;         /* 8 x R uint2 {index in tile list, 8x4 coverage mask}: per-block lists (ours) */
} gsr_binning_view;

typedef struct gsr_image_view {
  gsr_array_view accum_alpha; /* float per pixel: final transmittance T */
  gsr_array_view n_contrib;   /* uint32 per pixel */
  gsr_array_view ranges;      /* uint2 per tile */
  gsr_array_view tile_last;   /* uint32 per tile: max n_contrib over the tile's pixels (ours) */
  gsr_array_view tile_order;  /* uint32 per tile: launch order, longest lists first (ours) */
  gsr_array_view block_count; /* uint32 per 8x4 pixel block (8 per tile): candidates in the block's list (ours) */
  gsr_array_view last_group;  /* uint32 per pixel: 1 + group of 32 candidates of the pixel's last contributor (ours) */
  gsr_array_view summary;     /* 12 x uint32 per frame: {R, longest list, long lists, overflow, depth buckets, -,
                                 u64 pairs inside boxes (alphas evaluated), u64 pairs blended, -, -} (ours) */
} gsr_image_view;

GSR_API int gsr_abi_version(void);
/* sizeof of an ABI struct as this library was compiled: 0 gsr_fwd_args, 1 gsr_bwd_args, 2 gsr_array_view,
 * 3 gsr_geom_view, 4 gsr_binning_view, 5 gsr_image_view, 6 gsr_adam_group, 7 gsr_window_comm, 8 gsr_vgpr_params,
 * 9 gsr_async_status, 10 gsr_window_rows; 0 for an unknown index. Lets FFI bindings check layout. */
GSR_API size_t gsr_sizeof(int which);
GSR_API const char* gsr_last_error(void);

GSR_API size_t gsr_required_geom(int P);
GSR_API size_t gsr_required_image(int width, int height);
GSR_API size_t gsr_required_binning(int R);

GSR_API int gsr_forward(const gsr_fwd_args* args,
                gsr_alloc_fn geom_alloc, void* geom_ctx,
                gsr_alloc_fn binning_alloc, void* binning_ctx,
                gsr_alloc_fn image_alloc, void* image_ctx,
                void* stream, int* num_rendered);

/* Sync-free forward. The reference reads the number of instances R back to the host in the
 * middle of every forward (blocking cudaMemcpy, rasterizer_impl.cu:277) because its binning buffer is sized from
 * it. Here the caller provides all three buffers up front — geometry and image of gsr_required_geom/_image bytes,
 * binning of gsr_required_binning(R_capacity) bytes — and nothing waits for the GPU: every kernel reads what it
 * needs from device memory. `status` (pinned host memory, may be NULL) receives the counters through an
 * asynchronous copy on `stream`; it is valid once the caller has synchronised with the stream for its own
 * reasons (e.g. to read the loss). If status->overflow != 0 the frame did not fit (R > R_capacity; there is no
 * limit on the length of a tile's list) and the outputs are those of an EMPTY scene: redo the frame with
 * gsr_forward, or with a larger buffer. The matching gsr_backward takes R = R_capacity and the same three buffers. */
typedef struct gsr_async_status {
  uint32_t num_rendered; /* R */
  uint32_t longest_list;
  uint32_t long_lists;   /* tile lists cut into depth buckets before sorting (longer than the split_len knob) */
  uint32_t overflow;     /* 0, or 1: R > R_capacity */
} gsr_async_status;
GSR_API int gsr_forward_presized(const gsr_fwd_args* args, char* geom_buffer, char* binning_buffer, int R_capacity,
                                 char* image_buffer, void* stream, gsr_async_status* status);

GSR_API int gsr_backward(const gsr_bwd_args* args, void* stream);
/* Only the clearing of the four accumulated outputs (dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor: 44 B per
 * Gaussian) that gsr_backward() does first under GSR_BWD_OUTPUTS_UNINITIALIZED. It does not depend on the forward, so a
 * caller may issue it on a side stream while the forward runs and then pass GSR_BWD_ACCUMULATORS_ZEROED. Reads only
 * P and the four pointers of `args`. */
GSR_API int gsr_backward_prepare(const gsr_bwd_args* args, void* stream);

/* Gradient exchange of the view-sharded keyframe window (no counterpart in the reference, which is single-GPU;
 * SURVEY.md 8e): all-reduce (sum, FP32) of the flat per-Gaussian parameter-gradient buffer, written against peer
 * memory. The buffer is a symmetric allocation (same size on every rank, peer-mapped; e.g. torch's
 * _symmetric_memory). The caller brackets the calls with a barrier across the ranks on the same stream.
 *   gsr_window_allreduce_multimem: NVSwitch multicast path (multimem.ld_reduce + multimem.st); `multicast_buffer`
 *     is the multicast mapping of the buffer.
 *   gsr_window_allreduce_p2p: phase 1 = reduce-scatter, phase 2 = all-gather over plain peer loads, a barrier in
 *     between; peer_buffers[r] = rank r's buffer as mapped into this process (host array of `world` pointers).
 * n_floats must be a multiple of 4, buffers 16-byte aligned, world <= 16. */
GSR_API int gsr_window_allreduce_multimem(void* multicast_buffer, size_t n_floats, int rank, int world, void* stream);
GSR_API int gsr_window_allreduce_p2p(void* const* peer_buffers, size_t n_floats, int rank, int world, int phase,
                                     void* stream);

/* Sparse variant of the exchange. Floats [first_float, first_float + n_rows * floats_per_row) of the buffer are rows
 * of one Gaussian each (the 192-byte dL_dsh rows: 81 % of the gradient buffer) and row r is exchanged only if bit r
 * of union_words is set: the union over the ranks of "Gaussian r was visible in my view". A Gaussian visible in no
 * view of the window has an all-zero row on every rank, so the result is unchanged. first_float and floats_per_row
 * must be multiples of 4. rows == NULL or union_words == NULL: dense exchange.
 *   gsr_window_visibility_words: bit b of words[w] = radii[32 w + b] > 0, into THIS rank's part of a symmetric
 *     allocation (ceil(P / 32) words);
 *   gsr_window_union_words: out[w] = OR over ranks of peer_words[r][w] (after a barrier across the ranks);
 *   gsr_window_allreduce_{multimem,p2p}_rows: as above with the row mask. */
typedef struct gsr_window_rows {
  const uint32_t* union_words; /* device pointer, local */
  size_t first_float;
  uint32_t n_rows;
  uint32_t floats_per_row;
} gsr_window_rows;
GSR_API int gsr_window_visibility_words(const int* radii, int P, uint32_t* words, void* stream);
GSR_API int gsr_window_union_words(void* const* peer_words, int n_words, int world, uint32_t* out, void* stream);
GSR_API int gsr_window_allreduce_multimem_rows(void* multicast_buffer, size_t n_floats, int rank, int world,
                                               const gsr_window_rows* rows, void* stream);
GSR_API int gsr_window_allreduce_p2p_rows(void* const* peer_buffers, size_t n_floats, int rank, int world, int phase,
                                          const gsr_window_rows* rows, void* stream);

/* Factored exchange of the SH gradients (csrc/window_factored.cu; no counterpart in the reference). Per view the 48-float
 * dL_dsh row of a Gaussian is basis(dir_v) (x) dRGB_v (src/cuda_rasterizer/backward.cu:56-135), dir_v follows from the
 * replicated means and the view's camera centre, so the ranks exchange dRGB_v (3 floats per rendered Gaussian and view,
 * gathered) and all-reduce only the other 11 parameter-gradient floats; every rank then rebuilds the sum of the rows.
 *
 * Cross-rank ordering is inside the kernels. gsr_window_comm describes the ranks' flag pads: pads[r] = rank r's pad as
 * mapped into this process, gsr_window_pad_words() uint32 each, in peer-mapped (symmetric) memory, zero-initialised
 * once before the first step. Signals are release stores of the step's epoch, waits spin on the own pad; the epoch lives
 * in the own pad and is advanced by the step's last call, so a step can be captured in a CUDA graph and replayed.
 * The Gaussians may be cut into chunks that are pushed / reduced / expanded one after the other (`first`, `count`: the
 * chunk's Gaussians, first a multiple of 128; `sub` < 32: the chunk's number, ascending within a step), so that the
 * exchange of one chunk runs beside the per-Gaussian backward of the next. Every rank must issue the same sequence:
 *   gsr_window_push_rgb      (any stream; after the blend backward of this rank's view, GSR_BWD_ONLY_BLEND): writes the
 *       view's dRGB (dL_dcolor with the channels the forward clamped zeroed; culled Gaussians are skipped) and its
 *       visibility words into the view's slot of EVERY rank -- dst_rgb[r] / dst_vis[r], r < n_dst, are the slot as
 *       mapped for rank r (plain peer stores), or, with multicast != 0, dst_rgb[0] / dst_vis[0] its multicast mapping
 *       (multimem.st) -- then signals PUSHED(sub). A slot is gsr_window_slot_floats(P) floats (16-byte aligned) +
 *       ceil(P / 32) words. geom_buffer: of the view's forward.
 *   gsr_window_wait(GSR_WINDOW_CH_PUSHED, sub)   every peer's push of the chunk has landed here.
 *   gsr_window_expand_sh     dL_dsh[g] for the chunk's Gaussians = sum over v < n_views (in that order, identical on
 *       every rank) of basis(normalize(means3D - cam_pos[v])) (x) rgb_slots[v], skipping views whose visibility bit is
 *       clear; rows no view rendered are written as zeros. cam_pos: DEVICE array [n_views,3]; rgb_slots / vis_slots:
 *       host arrays of n_views device pointers; n_views <= 16, M == 16. row_state (optional): ceil(P / 32) words of
 *       device memory, all ones when dL_dsh holds arbitrary data, kept by the kernel from then on (bit = the row may
 *       be non-zero): rows no view rendered are then rewritten only if they still hold an earlier step's data. Chunks
 *       must start at multiples of 128 Gaussians for this (they do).
 *   gsr_window_allreduce_rest  (after the per-Gaussian backward of the chunk, GSR_BWD_ONLY_GEOM | GSR_BWD_SH_FACTORED):
 *       sum over the ranks of dL_dmean3D [P,3], dL_dscale [P,3], dL_drot [P,4], dL_dopacity [P] (the chunk's part), which
 *       start at floats first_float[0..3] (multiples of 4) of ONE symmetric buffer: buffers[r] = rank r's buffer
 *       (two-shot over peer loads), or with multicast != 0 buffers[0] = its multicast mapping (multimem.ld_reduce /
 *       multimem.st). With vis_slots (the gathered visibility words of the step's n_views views) groups of four
 *       Gaussians no view rendered are skipped; the kernel then waits for PUSHED(pushed_sub) itself (the pushes of a
 *       step are ordered, so the last chunk's signal covers them all). Waits for every rank's READY(sub) and signals
 *       DONE(sub).
 *   gsr_window_wait(GSR_WINDOW_CH_DONE, last sub, advance_epoch = 1)   LAST call of the step, after everything above:
 *       every rank's part of the all-reduce has landed here. */
#define GSR_WINDOW_FLAG_CHANNELS 4
#define GSR_WINDOW_CH_DONE 2
#define GSR_WINDOW_CH_PUSHED 3
typedef struct gsr_window_comm {
  int rank, world; /* world <= 16; world == 1: the calls do their local part only */
  void* pads[16];
} gsr_window_comm;
GSR_API size_t gsr_window_slot_floats(int P);
GSR_API size_t gsr_window_pad_words(void);
/* Synchronous (blocking copy): 0, or 1 + the channel on which a wait of this rank gave up after ~10 s (a peer died, or
 * the ranks issued different call sequences). The waits never hang the GPU; after a time-out the results are void. */
GSR_API int gsr_window_status(const gsr_window_comm* comm, uint32_t* timed_out_channel);
GSR_API int gsr_window_wait(const gsr_window_comm* comm, int channel, int sub, int advance_epoch, void* stream);
GSR_API int gsr_window_push_rgb(char* geom_buffer, int P, int first, int count, int sub, const int* radii,
                                const float* dL_dcolor, void* const* dst_rgb, void* const* dst_vis, int n_dst,
                                int multicast, const gsr_window_comm* comm, void* stream);
GSR_API int gsr_window_expand_sh(int P, int first, int count, int D, int M, int n_views, const float* means3D,
                                 const float* cam_pos, const void* const* rgb_slots, const void* const* vis_slots,
                                 float* dL_dsh, uint32_t* row_state, void* stream);
GSR_API int gsr_window_allreduce_rest(const gsr_window_comm* comm, void* const* buffers, int multicast, int P, int first,
                                      int count, int sub, int pushed_sub, const size_t* first_float, int n_views,
                                      const void* const* vis_slots, void* stream);

/* The two GPU steps right behind the rasterizer in GS-LIVM's mapping iteration, fused.
 *
 * gsr_image_loss <- gaussian_splatting::l1_loss + ssim and their combination (include/gs/gs/loss_utils.cuh:11-70,
 *   src/liw/lioOptimization.cpp:1729-1734): loss = (1 - lambda) mean|image - target| + lambda (1 - mean(SSIM map)), with
 *   the reference's own 11x11 window (loss_utils.cuh:24-30, NOT the textbook one) and zero padding, forward and backward
 *   in one call: dL_dimage [3,H,W] is what gsr_backward takes as dL_dpix. loss_out (device) receives
 *   {loss, mean|image - target|, mean(SSIM map)}; scratch = gsr_image_loss_scratch_floats(W, H) floats of device memory.
 *   lambda_dssim == 0: plain L1 in one pass.
 * gsr_adam_step <- torch::optim::Adam::step as GaussianModel::Training_setup sets it up (src/gs/gaussian.cu:396-428): no
 *   weight decay, no amsgrad; `groups` (host array, <= 8) are element ranges of the flat params / grads / exp_avg /
 *   exp_avg_sq buffers with their learning rates; `step` is the 1-based step count (bias correction). */
typedef struct gsr_adam_group {
  size_t first, count; /* element range [first, first + count) */
  float lr;
  /* period == 0: lr applies to the whole range. Otherwise the range is rows of `period` elements: the elements
   * [phase_lo, phase_hi) of every row take lr, the others lr_other. E.g. ONE [P,16,3] SH array instead of the
   * reference's _features_dc / _features_rest pair: {period 48, [0,3), lr_dc, lr_other = lr_dc / 20} (gaussian.cu:406-412). */
  unsigned period, phase_lo, phase_hi;
  float lr_other;
} gsr_adam_group;
GSR_API size_t gsr_image_loss_scratch_floats(int width, int height);
GSR_API int gsr_image_loss(const float* image, const float* target, int width, int height, float lambda_dssim,
                           float* dL_dimage, float* loss_out, float* scratch, void* stream);
GSR_API int gsr_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                          const gsr_adam_group* groups, int n_groups, float beta1, float beta2, float eps, int step,
                          void* stream);

/* VGPR + variance-centred Gaussian initialisation (csrc/vgpr.cu), the producer of the map's Gaussians:
 * gpProcess::forward_gp3d's per-voxel Gaussian-process regression (src/gp3d/gpprocess.cu:8-60,602-668) and
 * fastInitial3DGS (:420-458) + the scale rule of GaussianModel::Create_from_pcd (src/gs/gaussian.cu:325-386) in ONE
 * launch, one CTA per voxel, no cuBLAS, no per-frame allocation. Inputs per voxel: T <= 32 training points, the voxel's
 * integer coordinates and the regressed axis (the axis closest to the smallest-eigenvalue direction of the voxel's
 * points, cell.cpp:5-30; the caller buckets the sweep and runs the 3x3 PCA, gs_livm_b200/vgpr.py). Outputs: g*g
 * Gaussians per voxel, g = num_gp_side / neighbour: means [V*g*g,3], covariances [V*g*g,3,3], scales
 * sqrt(diag * scale_factor) [V*g*g,3]; var_mean [V] = 1 - mean of the variance entries (the reference's "needs more
 * points" gate, gpprocess.cu:63-122); optional test_points [V,side^2,3] and kvar [V,side^2] (parity tests). */
typedef struct gsr_vgpr_params {
  float grid;            /* voxel edge, config/basic_common.yaml:72 (0.2) */
  int num_gp_side;       /* lattice points per side (12), <= 16 */
  int neighbour;         /* lattice points per Gaussian and side (3); divides num_gp_side */
  float kernel_size;     /* k = exp(-kernel_size |d|) (1.0) */
  float variance_sensor; /* sigma: sigma^2 on Ky's diagonal (0.05) */
  float scale_factor;    /* (3.0) */
  int full_cover;        /* lattice from edge to edge instead of cell centres (0) */
} gsr_vgpr_params;
GSR_API int gsr_vgpr_gaussians(int V, int T, const float* train, const int* voxel, const int* direction,
                               const gsr_vgpr_params* params, float* means, float* covs, float* scales, float* var_mean,
                               float* test_points, float* kvar, void* stream);

GSR_API int gsr_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                     unsigned char* present, void* stream);

GSR_API int gsr_view_geom(char* geom_buffer, int P, gsr_geom_view* out);
GSR_API int gsr_view_binning(char* binning_buffer, int R, gsr_binning_view* out);
GSR_API int gsr_view_image(char* image_buffer, int width, int height, gsr_image_view* out);

/* Process-wide tuning / A-B knobs (no counterpart in the reference). Initial values come from the environment, read
 * ONCE on first use; gsr_configure() overrides them at any time (value < 0: back to the built-in default), takes
 * effect from the next forward / backward, and is thread-safe. Knobs (environment variable):
 *   "sort_device"   (GSR_SORT=device)   1: always bin with the reference's own sequence scan -> duplicateWithKeys ->
 *                                       device-wide radix sort -> ranges (parity tests of the intermediate arrays)
 *   "heavy_fwd", "heavy_bwd" (GSR_BLEND_HEAVY, both)  list length from which a tile's blend warps use four lanes per
 *                                       pixel; defaults 3072 / 4096
 *   "fwd_ilp" (GSR_FWD_ILP) 1|2|4, "bwd_ilp" (GSR_BWD_ILP) 1|2   pairs evaluated per lane per step
 *   "split_len" (GSR_SORT_SPLIT)        tile lists longer than this (rounded to 2^k - 1, 255..2047) are cut into depth
 *                                       buckets before sorting; "bucket_target" (GSR_SORT_BUCKET) instances per bucket
 *   "ar_grid" (GSR_AR_GRID)             CTAs per SM of the window exchange kernels
 * Returns GSR_ERR_INVALID for an unknown knob. */
GSR_API int gsr_configure(const char* knob, int value);
GSR_API int gsr_config_get(const char* knob, int* value);

/* Optional per-stage device timing (CUDA events recorded on the launch stream around each stage of the
 * last forward/backward on the calling thread); used by bench.py for the roofline figures. */
enum gsr_stage {
  GSR_STAGE_PREPROCESS = 0, /* preprocess_fwd            <- preprocessCUDA fwd        (forward.cu:179-286)   */
  GSR_STAGE_SCAN = 1,       /* prefix sum                <- cub InclusiveSum          (rasterizer_impl.cu:270) */
  GSR_STAGE_DUPLICATE = 2,  /* duplicate_with_keys       <- duplicateWithKeys         (rasterizer_impl.cu:64)  */
  GSR_STAGE_SORT = 3,       /* (tile|depth) radix sort   <- cub SortPairs             (rasterizer_impl.cu:298) */
  GSR_STAGE_FINALIZE = 4,   /* per-block candidate lists (block_lists.cu); device-wide path: + identifyTileRanges */
  GSR_STAGE_BLEND_FWD = 5,  /* blend_fwd                 <- renderCUDA fwd            (forward.cu:291-407)   */
  GSR_STAGE_BLEND_BWD = 6,  /* blend_bwd                 <- renderCUDA bwd            (backward.cu:438-603)  */
  GSR_STAGE_GEOM_BWD = 7,   /* fused per-Gaussian bwd    <- computeCov2DCUDA + preprocessCUDA bwd (backward.cu:140-435) */
  GSR_NUM_STAGES = 8
};
GSR_API void gsr_stage_timing_enable(int on);
GSR_API int gsr_stage_times(float* ms_out, int n); /* ms per stage, -1 where the stage did not run */

/* Stream used by the C++ shim (CudaRasterizer::Rasterizer has no stream parameter). Default NULL =
 * legacy default stream, identical to the reference's launches. Thread-local. */
GSR_API void gsr_shim_set_stream(void* stream);
GSR_API void* gsr_shim_get_stream(void);

#ifdef __cplusplus
}
#endif
#endif /* GSR_B200_H_INCLUDED */
