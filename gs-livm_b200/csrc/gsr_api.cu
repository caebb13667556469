// C ABI of libgsr_b200 (declared in include/gsr.h): scratch-state carving, stage orchestration and
// error handling for one forward / backward. Mirrors the control flow of the reference's
// CudaRasterizer::Rasterizer::forward/backward (src/cuda_rasterizer/rasterizer_impl.cu:181-457)
// with our kernels; nothing here touches the CPU oracle and there is no fallback path.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "gsr_common.cuh"

namespace gsr {

static thread_local char t_err[512] = "";
static thread_local void* t_shim_stream = nullptr;

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof(t_err), fmt, ap);
  va_end(ap);
  return code;
}

// ---- optional per-stage CUDA-event timing (bench.py's roofline numbers) ---------------------------
// When enabled, an event is recorded on the launch stream before/after every stage of the last
// forward and backward; gsr_stage_times() synchronises on them and returns milliseconds.
struct StageTimer {
  bool enabled = false;
  bool created = false;
  bool seen[GSR_NUM_STAGES] = {};
  cudaEvent_t beg[GSR_NUM_STAGES], end[GSR_NUM_STAGES];
};
static thread_local StageTimer t_timer;

struct StageScope {
  int id;
  cudaStream_t s;
  bool on;
  StageScope(int id_, cudaStream_t s_) : id(id_), s(s_), on(t_timer.enabled) {
    if (!on) return;
    if (!t_timer.created) {
      for (int i = 0; i < GSR_NUM_STAGES; ++i) cudaEventCreate(&t_timer.beg[i]), cudaEventCreate(&t_timer.end[i]);
      t_timer.created = true;
    }
    cudaEventRecord(t_timer.beg[id], s);
  }
  ~StageScope() {
    if (!on) return;
    cudaEventRecord(t_timer.end[id], s);
    t_timer.seen[id] = true;
  }
};

#define GSR_CUDA(expr, what)                                                                      \
  do {                                                                                            \
    cudaError_t e__ = (expr);                                                                     \
    if (e__ != cudaSuccess) return fail(GSR_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e__));   \
  } while (0)

// debug == the reference's CHECK_CUDA(A, debug): synchronise and surface asynchronous errors per stage
#define GSR_STAGE(id, expr, what, debug, stream)                                                  \
  do {                                                                                            \
    {                                                                                             \
      StageScope scope__(id, stream);                                                             \
      GSR_CUDA((expr), what);                                                                     \
    }                                                                                             \
    if (debug) GSR_CUDA(cudaStreamSynchronize(stream), what);                                     \
  } while (0)

// ---- tuning knobs: atomics initialised once from the environment, changed with gsr_configure() ---------------
namespace {
struct Knob {
  const char* name;
  const char* env;
  int def, lo, hi;
  std::atomic<int> value;
};
enum { K_SORT_DEVICE, K_HEAVY_FWD, K_HEAVY_BWD, K_FWD_ILP, K_BWD_ILP, K_SPLIT, K_BUCKET, K_AR_GRID, K_COUNT };
Knob g_knobs[K_COUNT] = {
    {"sort_device", "GSR_SORT", 0, 0, 1, {0}},          {"heavy_fwd", "GSR_BLEND_HEAVY", 3072, 1, 1 << 30, {3072}},
    {"heavy_bwd", "GSR_BLEND_HEAVY", 4096, 1, 1 << 30, {4096}}, {"fwd_ilp", "GSR_FWD_ILP", 4, 1, 4, {4}},
    {"bwd_ilp", "GSR_BWD_ILP", 2, 1, 2, {2}},            {"split_len", "GSR_SORT_SPLIT", 2047, 255, kSegCap - 1, {2047}},
    {"bucket_target", "GSR_SORT_BUCKET", 640, kMinBucketTarget, kSegCap, {640}},
    {"ar_grid", "GSR_AR_GRID", 2, 1, 16, {2}},
};
std::once_flag g_knobs_once;

int sanitize(int k, int v) {
  v = v < g_knobs[k].lo ? g_knobs[k].lo : (v > g_knobs[k].hi ? g_knobs[k].hi : v);
  if (k == K_SPLIT) {  // 2^j - 1: the long-list class must be a prefix of tile_order's log2 buckets
    int p = 256;
    while (p * 2 - 1 <= v) p *= 2;
    v = p - 1;
  }
  return v;
}

void knobs_from_env() {
  for (int k = 0; k < K_COUNT; ++k) {
    const char* e = getenv(g_knobs[k].env);
    if (!e || !*e) continue;
    int v;
    if (k == K_SORT_DEVICE) v = !strcmp(e, "device") ? 1 : 0;
    else v = atoi(e);
    g_knobs[k].value.store(sanitize(k, v), std::memory_order_relaxed);
  }
}
int knob(int k) {
  std::call_once(g_knobs_once, knobs_from_env);
  return g_knobs[k].value.load(std::memory_order_relaxed);
}
}  // namespace

Tuning tuning() {
  Tuning t;
  t.sort_device = knob(K_SORT_DEVICE);
  t.heavy_fwd = knob(K_HEAVY_FWD), t.heavy_bwd = knob(K_HEAVY_BWD), t.fwd_ilp = knob(K_FWD_ILP), t.bwd_ilp = knob(K_BWD_ILP);
  t.split_len = knob(K_SPLIT), t.bucket_target = knob(K_BUCKET), t.ar_grid = knob(K_AR_GRID);
  return t;
}

// ---- per-device caches, one mutex (the forward runs on the mapping thread, the backward on an autograd worker) -----
namespace {
std::mutex g_dev_mutex;
int g_sm_count[64] = {};
struct SmemAttr {
  const void* kernel;
  int dev, bytes;
};
std::vector<SmemAttr> g_smem_attrs;
}  // namespace

int device_sm_count() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  std::lock_guard<std::mutex> lock(g_dev_mutex);
  if (!g_sm_count[dev] && cudaDeviceGetAttribute(&g_sm_count[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    g_sm_count[dev] = 148;
  return g_sm_count[dev];
}

cudaError_t ensure_dynamic_smem(const void* kernel, int bytes) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(g_dev_mutex);
  for (const SmemAttr& a : g_smem_attrs)
    if (a.kernel == kernel && a.dev == dev && a.bytes >= bytes) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) g_smem_attrs.push_back({kernel, dev, bytes});
  return e;
}

GeomState carve_geom(char* chunk, int P) {
  GeomState g;
  carve(chunk, g.rec, (size_t)P);
  carve(chunk, g.tiles_touched, (size_t)P);
  carve(chunk, g.point_offsets, (size_t)P);
  carve(chunk, g.internal_radii, (size_t)P);
  carve(chunk, g.clamped, (size_t)P);
  g.scan_temp_bytes = scan_temp_bytes(P);
  carve(chunk, g.scan_temp, g.scan_temp_bytes);
  return g;
}

ImageState carve_image(char* chunk, int width, int height) {
  ImageState s;
  const size_t N = (size_t)width * height;
  const size_t tiles = (size_t)((width + kTile - 1) / kTile) * ((height + kTile - 1) / kTile);
  carve(chunk, s.final_T, N);
  carve(chunk, s.n_contrib, N);
  carve(chunk, s.ranges, tiles);
  carve(chunk, s.tile_last, tiles);
  carve(chunk, s.tile_order, tiles);
  carve(chunk, s.tile_count, tiles);
  carve(chunk, s.bucket_base, tiles);
  carve(chunk, s.summary, 12);
  carve(chunk, s.block_count, tiles * 8);
  carve(chunk, s.bdesc, tiles * 8);
  carve(chunk, s.last_group, N);
  carve(chunk, s.tail_hits, tiles * 8 * 32);
  return s;
}

BinningState carve_binning(char* chunk, int R) {
  BinningState b;
  const size_t n = (size_t)R;
  carve(chunk, b.inst, n);
  carve(chunk, b.keys_unsorted, n);
  carve(chunk, b.keys, n);
  carve(chunk, b.vals_unsorted, n);
  carve(chunk, b.vals, n + 4);  // +16 B: blend_bwd bulk-copies ids from an aligned-down start
  carve(chunk, b.vals_tmp, n);
  carve(chunk, b.seg_desc, max_depth_buckets(n));
  carve(chunk, b.blist, 8 * n);
  carve(chunk, b.bhits, 8 * n);
  b.sort_temp_bytes = sort_temp_bytes(R);
  carve(chunk, b.sort_temp, b.sort_temp_bytes);
  return b;
}

template <typename F>
static size_t required_bytes(F carve_fn) {
  // carve from address 0 and read back the end, then add one alignment unit because the caller's
  // buffer may start anywhere (same trick as the reference's required<T>, rasterizer_impl.h:61-66)
  return carve_fn() + kAlign;
}

// Eight pinned words per host thread for the {R, longest list, #long lists, -, #depth buckets} read-back (reference:
// blocking cudaMemcpy D2H, rasterizer_impl.cu:277).
static uint32_t* pinned_summary() {
  static thread_local uint32_t* p = nullptr;
  if (!p) {
    if (cudaMallocHost(reinterpret_cast<void**>(&p), 8 * sizeof(uint32_t)) != cudaSuccess) p = nullptr;
  }
  return p;
}

}  // namespace gsr

using namespace gsr;

extern "C" {

int gsr_abi_version(void) { return GSR_ABI_VERSION; }
const char* gsr_last_error(void) { return t_err; }

size_t gsr_sizeof(int which) {
  switch (which) {
    case 0: return sizeof(gsr_fwd_args);
    case 1: return sizeof(gsr_bwd_args);
    case 2: return sizeof(gsr_array_view);
    case 3: return sizeof(gsr_geom_view);
    case 4: return sizeof(gsr_binning_view);
    case 5: return sizeof(gsr_image_view);
    case 6: return sizeof(gsr_adam_group);
    case 7: return sizeof(gsr_window_comm);
    case 8: return sizeof(gsr_vgpr_params);
    case 9: return sizeof(gsr_async_status);
    case 10: return sizeof(gsr_window_rows);
    default: return 0;
  }
}

int gsr_configure(const char* name, int value) {
  if (!name) return fail(GSR_ERR_INVALID, "gsr_configure: null knob name");
  for (int k = 0; k < K_COUNT; ++k)
    if (!strcmp(name, g_knobs[k].name)) {
      (void)knob(k);  // the environment is read first, once; an explicit setting then wins
      g_knobs[k].value.store(value < 0 ? g_knobs[k].def : sanitize(k, value), std::memory_order_relaxed);
      return GSR_OK;
    }
  return fail(GSR_ERR_INVALID, "gsr_configure: unknown knob '%s'", name);
}

int gsr_config_get(const char* name, int* value) {
  if (!name || !value) return fail(GSR_ERR_INVALID, "gsr_config_get: null argument");
  for (int k = 0; k < K_COUNT; ++k)
    if (!strcmp(name, g_knobs[k].name)) {
      *value = knob(k);
      return GSR_OK;
    }
  return fail(GSR_ERR_INVALID, "gsr_config_get: unknown knob '%s'", name);
}

void gsr_stage_timing_enable(int on) { t_timer.enabled = on != 0; }

int gsr_stage_times(float* ms, int n) {
  if (!ms || n < GSR_NUM_STAGES) return fail(GSR_ERR_INVALID, "gsr_stage_times: need room for %d floats", GSR_NUM_STAGES);
  for (int i = 0; i < GSR_NUM_STAGES; ++i) {
    ms[i] = -1.0f;
    if (!t_timer.created || !t_timer.seen[i]) continue;
    GSR_CUDA(cudaEventSynchronize(t_timer.end[i]), "stage timer");
    GSR_CUDA(cudaEventElapsedTime(&ms[i], t_timer.beg[i], t_timer.end[i]), "stage timer");
  }
  return GSR_OK;
}

void gsr_shim_set_stream(void* stream) { t_shim_stream = stream; }
void* gsr_shim_get_stream(void) { return t_shim_stream; }

size_t gsr_required_geom(int P) {
  if (P < 0) return 0;
  return required_bytes([&] {
    GeomState g = carve_geom(nullptr, P);
    return (size_t)(reinterpret_cast<uintptr_t>(g.scan_temp) + g.scan_temp_bytes);
  });
}

size_t gsr_required_image(int width, int height) {
  if (width < 0 || height < 0) return 0;
  return required_bytes([&] {
    ImageState s = carve_image(nullptr, width, height);
    const size_t tiles = (size_t)((width + kTile - 1) / kTile) * ((height + kTile - 1) / kTile);
    return (size_t)(reinterpret_cast<uintptr_t>(s.tail_hits + tiles * 8 * 32));
  });
}

size_t gsr_required_binning(int R) {
  if (R < 0) return 0;
  return required_bytes([&] {
    BinningState b = carve_binning(nullptr, R);
    return (size_t)(reinterpret_cast<uintptr_t>(b.sort_temp) + b.sort_temp_bytes);
  });
}

static int validate_fwd(const gsr_fwd_args& a, const char* who) {
  if (a.P <= 0 || a.width <= 0 || a.height <= 0) return fail(GSR_ERR_INVALID, "%s: P, width, height must be > 0", who);
  if (!a.means3D || !a.opacities || !a.viewmatrix || !a.projmatrix || !a.cam_pos || !a.background || !a.out_color ||
      !a.out_depth || !a.out_acc)
    return fail(GSR_ERR_INVALID, "%s: null input/output pointer", who);
  if (!a.colors_precomp && (!a.shs || a.M <= 0)) return fail(GSR_ERR_INVALID, "%s: need shs or colors_precomp", who);
  if (!a.colors_precomp && (a.D < 0 || a.D > 3 || (a.D + 1) * (a.D + 1) > a.M))
    return fail(GSR_ERR_INVALID, "%s: SH degree %d needs %d coefficients, M = %d", who, a.D, (a.D + 1) * (a.D + 1), a.M);
  if (!a.cov3D_precomp && (!a.scales || !a.rotations))
    return fail(GSR_ERR_INVALID, "%s: need scales+rotations or cov3D_precomp", who);
  return GSR_OK;
}

static int launch_blend_forward(const gsr_fwd_args& a, const ImageState& img, const BinningState& bin, int tiles,
                                int grid_x, int debug, cudaStream_t stream) {
  GSR_STAGE(GSR_STAGE_FINALIZE, launch_block_lists(tiles, img, bin, grid_x, stream), "block_lists", debug, stream);
  GSR_STAGE(GSR_STAGE_BLEND_FWD,
            launch_blend_fwd_sparse(a.width, a.height, img, bin, a.background, a.out_color, a.out_depth, a.out_acc,
                                    stream),
            "blend_fwd", debug, stream);
  return GSR_OK;
}

int gsr_forward(const gsr_fwd_args* args, gsr_alloc_fn geom_alloc, void* geom_ctx, gsr_alloc_fn binning_alloc,
                void* binning_ctx, gsr_alloc_fn image_alloc, void* image_ctx, void* stream_v, int* num_rendered) {
  if (!args || !geom_alloc || !binning_alloc || !image_alloc || !num_rendered)
    return fail(GSR_ERR_INVALID, "gsr_forward: null argument");
  const gsr_fwd_args& a = *args;
  *num_rendered = 0;
  if (int st = validate_fwd(a, "gsr_forward")) return st;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  const int debug = a.debug;

  const float focal_y = a.height / (2.0f * a.tan_fovy);  // rasterizer_impl.cu:210-211
  const float focal_x = a.width / (2.0f * a.tan_fovx);
  const int grid_x = (a.width + kTile - 1) / kTile, grid_y = (a.height + kTile - 1) / kTile;
  const int tiles = grid_x * grid_y;

  char* geom_chunk = geom_alloc(geom_ctx, gsr_required_geom(a.P));
  if (!geom_chunk) return fail(GSR_ERR_ALLOC, "gsr_forward: geometry allocator returned NULL");
  GeomState geom = carve_geom(geom_chunk, a.P);
  int* radii = a.radii ? a.radii : geom.internal_radii;

  char* img_chunk = image_alloc(image_ctx, gsr_required_image(a.width, a.height));
  if (!img_chunk) return fail(GSR_ERR_ALLOC, "gsr_forward: image allocator returned NULL");
  ImageState img = carve_image(img_chunk, a.width, a.height);

  // instances per tile are counted by preprocess; one small kernel turns them into the tile ranges, the
  // scatter cursors, the launch order of the per-tile CTAs and {R, longest list, #long lists}
  GSR_CUDA(cudaMemsetAsync(img.tile_count, 0, (size_t)tiles * sizeof(uint32_t), stream), "clear tile counts");
  GSR_STAGE(GSR_STAGE_PREPROCESS, launch_preprocess_fwd(a, geom, radii, img.tile_count, focal_x, focal_y, stream),
            "preprocess_fwd", debug, stream);
  GSR_STAGE(GSR_STAGE_SCAN, launch_tile_scan(tiles, img, 0xffffffffu, stream), "tile_scan", debug, stream);

  // R has to reach the host: the binning allocator is a host callback (reference: blocking cudaMemcpy D2H,
  // rasterizer_impl.cu:277)
  uint32_t* h_summary = pinned_summary();
  if (!h_summary) return fail(GSR_ERR_CUDA, "gsr_forward: cudaMallocHost failed");
  GSR_CUDA(cudaMemcpyAsync(h_summary, img.summary, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream),
           "read num_rendered");
  GSR_CUDA(cudaStreamSynchronize(stream), "read num_rendered");
  if (h_summary[0] > 0x7fffffffu) return fail(GSR_ERR_INVALID, "gsr_forward: instance count overflows int32");
  const int R = (int)h_summary[0];
  const int n_long = (int)h_summary[2], n_buckets = (int)h_summary[4];

  char* bin_chunk = binning_alloc(binning_ctx, gsr_required_binning(R));
  if (!bin_chunk) return fail(GSR_ERR_ALLOC, "gsr_forward: binning allocator returned NULL");
  BinningState bin = carve_binning(bin_chunk, R);

  if (R > 0) {
    if (!tuning().sort_device) {
      // per-tile path: bucket by tile, long lists cut into depth buckets, every segment sorted inside one CTA
      // (tile_sort.cu); any list length
      GSR_STAGE(GSR_STAGE_DUPLICATE, launch_scatter_instances(a.P, geom, radii, img, bin, grid_x, grid_y, stream),
                "scatter_instances", debug, stream);
      GSR_STAGE(GSR_STAGE_SORT, launch_tile_sort(tiles, n_long, n_buckets, R, img, bin, geom, grid_x, stream),
                "tile_sort", debug, stream);
    } else {
      // (parity tests only: GSR_SORT=device / gsr_configure("sort_device", 1))
      // device-wide path, the reference's own sequence: scan -> duplicateWithKeys -> radix sort -> ranges
      GSR_STAGE(GSR_STAGE_DUPLICATE,
                [&] {
                  cudaError_t e = run_scan(geom, a.P, stream);
                  return e != cudaSuccess ? e : launch_duplicate_with_keys(a.P, geom, radii, bin, grid_x, grid_y, stream);
                }(),
                "duplicate_with_keys", debug, stream);
      const int bit = (int)higher_msb((uint32_t)tiles);
      GSR_STAGE(GSR_STAGE_SORT, run_sort(bin, R, 32 + bit, stream), "sort", debug, stream);
      // no stage timer here: on this path GSR_STAGE_FINALIZE times block_lists (launch_blend_forward)
      GSR_CUDA(launch_finalize_binning(R, bin, geom, grid_x, img.ranges, stream), "finalize_binning");
      if (debug) GSR_CUDA(cudaStreamSynchronize(stream), "finalize_binning");
    }
  }
  if (int st = launch_blend_forward(a, img, bin, tiles, grid_x, debug, stream)) return st;
  *num_rendered = R;
  return GSR_OK;
}

int gsr_forward_presized(const gsr_fwd_args* args, char* geom_buffer, char* binning_buffer, int R_capacity,
                         char* image_buffer, void* stream_v, gsr_async_status* status) {
  if (!args || !geom_buffer || !binning_buffer || !image_buffer || R_capacity <= 0)
    return fail(GSR_ERR_INVALID, "gsr_forward_presized: null buffer or non-positive capacity");
  const gsr_fwd_args& a = *args;
  if (int st = validate_fwd(a, "gsr_forward_presized")) return st;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  const int debug = a.debug;
  const float focal_y = a.height / (2.0f * a.tan_fovy);
  const float focal_x = a.width / (2.0f * a.tan_fovx);
  const int grid_x = (a.width + kTile - 1) / kTile, grid_y = (a.height + kTile - 1) / kTile;
  const int tiles = grid_x * grid_y;
  GeomState geom = carve_geom(geom_buffer, a.P);
  int* radii = a.radii ? a.radii : geom.internal_radii;
  ImageState img = carve_image(image_buffer, a.width, a.height);
  BinningState bin = carve_binning(binning_buffer, R_capacity);

  GSR_CUDA(cudaMemsetAsync(img.tile_count, 0, (size_t)tiles * sizeof(uint32_t), stream), "clear tile counts");
  GSR_STAGE(GSR_STAGE_PREPROCESS, launch_preprocess_fwd(a, geom, radii, img.tile_count, focal_x, focal_y, stream),
            "preprocess_fwd", debug, stream);
  // tile_scan empties every range if the frame does not fit, so the stages below then find nothing to do
  GSR_STAGE(GSR_STAGE_SCAN, launch_tile_scan(tiles, img, (uint32_t)R_capacity, stream), "tile_scan", debug, stream);
  if (status)
    GSR_CUDA(cudaMemcpyAsync(status, img.summary, sizeof(gsr_async_status), cudaMemcpyDeviceToHost, stream),
             "copy status");
  GSR_STAGE(GSR_STAGE_DUPLICATE, launch_scatter_instances(a.P, geom, radii, img, bin, grid_x, grid_y, stream),
            "scatter_instances", debug, stream);
  GSR_STAGE(GSR_STAGE_SORT, launch_tile_sort(tiles, -1, -1, R_capacity, img, bin, geom, grid_x, stream), "tile_sort",
            debug, stream);
  return launch_blend_forward(a, img, bin, tiles, grid_x, debug, stream);
}

// The four outputs blend_bwd accumulates into with atomics start from zero; adjacent arrays (one flat gradient buffer,
// as gs_livm_b200.engine lays them out) are cleared by one memset.
static int zero_accumulators(const gsr_bwd_args& a, cudaStream_t stream) {
  struct Span {
    char* p;
    size_t n;
  } spans[4] = {{reinterpret_cast<char*>(a.dL_dmean2D), 3 * (size_t)a.P * sizeof(float)},
                {reinterpret_cast<char*>(a.dL_dconic), 4 * (size_t)a.P * sizeof(float)},
                {reinterpret_cast<char*>(a.dL_dopacity), (size_t)a.P * sizeof(float)},
                {reinterpret_cast<char*>(a.dL_dcolor), 3 * (size_t)a.P * sizeof(float)}};
  for (int i = 0; i < 4; ++i)
    for (int j = i + 1; j < 4; ++j)
      if (spans[j].p < spans[i].p) {
        Span t = spans[i];
        spans[i] = spans[j], spans[j] = t;
      }
  for (int i = 0; i < 4;) {
    char* p = spans[i].p;
    size_t n = spans[i].n;
    int j = i + 1;
    while (j < 4 && spans[j].p == p + n) n += spans[j++].n;
    GSR_CUDA(cudaMemsetAsync(p, 0, n, stream), "zero accumulated gradients");
    i = j;
  }
  return GSR_OK;
}

int gsr_backward_prepare(const gsr_bwd_args* args, void* stream_v) {
  if (!args || args->P <= 0 || !args->dL_dmean2D || !args->dL_dconic || !args->dL_dopacity || !args->dL_dcolor)
    return fail(GSR_ERR_INVALID, "gsr_backward_prepare: bad argument");
  return zero_accumulators(*args, static_cast<cudaStream_t>(stream_v));
}

int gsr_backward(const gsr_bwd_args* args, void* stream_v) {
  if (!args) return fail(GSR_ERR_INVALID, "gsr_backward: null argument");
  const gsr_bwd_args& a = *args;
  if (a.P <= 0 || a.width <= 0 || a.height <= 0 || a.R < 0) return fail(GSR_ERR_INVALID, "gsr_backward: bad sizes");
  if (!a.geom_buffer || !a.binning_buffer || !a.image_buffer) return fail(GSR_ERR_INVALID, "gsr_backward: null state buffer");
  if (!a.means3D || !a.viewmatrix || !a.projmatrix || !a.cam_pos || !a.background || !a.dL_dpix || !a.dL_dacc ||
      !a.dL_dmean2D || !a.dL_dconic || !a.dL_dopacity || !a.dL_dcolor || !a.dL_dmean3D || !a.dL_dcov3D)
    return fail(GSR_ERR_INVALID, "gsr_backward: null input/output pointer");
  if (!a.colors_precomp && (!a.shs || (!a.dL_dsh && !(a.flags & GSR_BWD_SH_FACTORED))))
    return fail(GSR_ERR_INVALID, "gsr_backward: need shs + dL_dsh");
  if ((a.flags & GSR_BWD_ONLY_BLEND) && (a.flags & GSR_BWD_ONLY_GEOM))
    return fail(GSR_ERR_INVALID, "gsr_backward: GSR_BWD_ONLY_BLEND and GSR_BWD_ONLY_GEOM exclude each other");
  if (!a.cov3D_precomp && (!a.scales || !a.rotations || !a.dL_dscale || !a.dL_drot))
    return fail(GSR_ERR_INVALID, "gsr_backward: need scales/rotations and their gradient outputs");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  const int debug = a.debug;

  GeomState geom = carve_geom(a.geom_buffer, a.P);
  BinningState bin = carve_binning(a.binning_buffer, a.R);
  ImageState img = carve_image(a.image_buffer, a.width, a.height);
  const int* radii = a.radii ? a.radii : geom.internal_radii;
  const float focal_y = a.height / (2.0f * a.tan_fovy);
  const float focal_x = a.width / (2.0f * a.tan_fovx);

  const bool do_blend = !(a.flags & GSR_BWD_ONLY_GEOM), do_geom = !(a.flags & GSR_BWD_ONLY_BLEND);
  if (do_blend && (a.flags & GSR_BWD_OUTPUTS_UNINITIALIZED) && !(a.flags & GSR_BWD_ACCUMULATORS_ZEROED))
    if (int st = zero_accumulators(a, stream)) return st;
  if (do_blend && a.R > 0) {
    GSR_STAGE(GSR_STAGE_BLEND_BWD,
              launch_blend_bwd_sparse(a.width, a.height, img, bin, a.background, a.dL_dpix, a.dL_dacc, a.dL_dmean2D,
                                      a.dL_dconic, a.dL_dopacity, a.dL_dcolor, stream),
              "blend_bwd", debug, stream);
  }
  if (do_geom)
    GSR_STAGE(GSR_STAGE_GEOM_BWD, launch_geom_bwd(a, geom, radii, focal_x, focal_y, stream), "geom_bwd", debug, stream);
  return GSR_OK;
}

int gsr_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* /*projmatrix*/,
                     unsigned char* present, void* stream) {
  if (P < 0 || (P > 0 && (!means3D || !viewmatrix || !present))) return fail(GSR_ERR_INVALID, "gsr_mark_visible: bad argument");
  if (P == 0) return GSR_OK;
  GSR_CUDA(launch_mark_visible(P, means3D, viewmatrix, present, static_cast<cudaStream_t>(stream)), "mark_visible");
  return GSR_OK;
}

static gsr_array_view view_of(void* p, size_t stride, size_t count) {
  gsr_array_view v;
  v.ptr = p, v.stride = stride, v.count = count;
  return v;
}

int gsr_view_geom(char* geom_buffer, int P, gsr_geom_view* out) {
  if (!geom_buffer || P < 0 || !out) return fail(GSR_ERR_INVALID, "gsr_view_geom: bad argument");
  GeomState g = carve_geom(geom_buffer, P);
  char* r = reinterpret_cast<char*>(g.rec);
  const size_t n = (size_t)P, st = sizeof(SplatRecord);
  out->means2D = view_of(r + 0, st, n);
  out->extent = view_of(r + 8, st, n);
  out->conic_opacity = view_of(r + 16, st, n);
  out->rgb = view_of(r + 32, st, n);
  out->depths = view_of(r + 44, st, n);
  out->clamped = view_of(g.clamped, 1, n);
  out->internal_radii = view_of(g.internal_radii, 4, n);
  out->tiles_touched = view_of(g.tiles_touched, 4, n);
  out->point_offsets = view_of(g.point_offsets, 4, n);
  return GSR_OK;
}

int gsr_view_binning(char* binning_buffer, int R, gsr_binning_view* out) {
  if (!binning_buffer || R < 0 || !out) return fail(GSR_ERR_INVALID, "gsr_view_binning: bad argument");
  BinningState b = carve_binning(binning_buffer, R);
  const size_t n = (size_t)R;
  out->point_list_keys_unsorted = view_of(b.keys_unsorted, 8, n);
  out->point_list_unsorted = view_of(b.vals_unsorted, 4, n);
  out->point_list_keys = view_of(b.keys, 8, n);
  out->point_list = view_of(b.vals, 4, n);
  out->instance_records = view_of(b.inst, sizeof(SplatRecord), n);
  out->block_candidates = view_of(b.blist, sizeof(uint2), 8 * n);
  return GSR_OK;
}

int gsr_view_image(char* image_buffer, int width, int height, gsr_image_view* out) {
  if (!image_buffer || width <= 0 || height <= 0 || !out) return fail(GSR_ERR_INVALID, "gsr_view_image: bad argument");
  ImageState s = carve_image(image_buffer, width, height);
  const size_t N = (size_t)width * height;
  const size_t tiles = (size_t)((width + kTile - 1) / kTile) * ((height + kTile - 1) / kTile);
  out->accum_alpha = view_of(s.final_T, 4, N);
  out->n_contrib = view_of(s.n_contrib, 4, N);
  out->ranges = view_of(s.ranges, 8, tiles);
  out->tile_last = view_of(s.tile_last, 4, tiles);
  out->tile_order = view_of(s.tile_order, 4, tiles);
  out->block_count = view_of(s.block_count, 4, tiles * 8);
  out->last_group = view_of(s.last_group, 4, N);
  out->summary = view_of(s.summary, 4, 12);
  return GSR_OK;
}

}  // extern "C"
