// wx_api.cu -- the C ABI of include/woxel_b200.h: context, tree packing/upload, frame entry points.
//
// Replaces the wgpu plumbing of the reference around its compute pass: device creation
// (src/render/wgpu_context.rs:33-99), model upload (:101-159, :506-573) and the per-frame encode
// + dispatch (:207-292).  Unlike the reference, nothing is re-created per frame (:219-268 builds
// pipelines, layouts and textures on every redraw): buffers, streams and events live in the context.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <memory>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include <nvtx3/nvToolsExt.h>  // header-only; a no-op unless a profiler injects the NVTX library

#include "wx_device.cuh"
#include "wx_internal.h"
#include "wx_pack.h"

using namespace wx;

// Every entry point that allocates on the host (vectors, threads) is a function-try-block: no C++ exception crosses the C ABI
// (bad_alloc -> WX_ERR_OUT_OF_MEMORY, anything else -> WX_ERR_UNSUPPORTED with the text in wx_last_error(NULL)).

// ---------------------------------------------------------------------------------------------
// Context
// ---------------------------------------------------------------------------------------------

// Long-tiles-first state (wx_internal.h TileSched) of ONE launch geometry on ONE stream: two sets of (list, flags), the
// previous launch's and the one being recorded, swapped after every launch; a high-priority stream for the long-tile kernel.
// Everything a launch touches is ordered on the launch's stream, so launches of the same geometry on the same stream never
// race; other streams / geometries get their own entry.
struct SchedKey {
  const void* tree;
  const void* stream;
  uint32_t width, height, cam0, ncam, shard_index, shard_count, band_rows, row0, row1;
  bool operator==(const SchedKey& o) const { return memcmp(this, &o, sizeof(*this)) == 0; }
};
struct SchedEntry final : TileSched {
  SchedKey key{};
  uint32_t* list[2] = {nullptr, nullptr};
  uint32_t* flag[2] = {nullptr, nullptr};
  uint32_t n_tiles = 0, cap = 0;
  int cur = 0;
  bool valid = false;  // set `cur` describes a finished launch of this geometry
  cudaStream_t ls = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
  uint64_t last_use = 0;

  void release() {
    if (ls) (void)cudaStreamSynchronize(ls);
    for (int k = 0; k < 2; ++k) {
      if (flag[k]) (void)cudaFree(flag[k]);  // (list[k] points into the same allocation)
      list[k] = flag[k] = nullptr;
    }
    n_tiles = cap = 0, valid = false, prepared = false;
  }
  ~SchedEntry() override {
    release();
    if (ls) (void)cudaStreamDestroy(ls);
    if (fork) (void)cudaEventDestroy(fork);
    if (join) (void)cudaEventDestroy(join);
  }
  cudaError_t prepare(RenderParams& P, uint32_t n_cams, uint32_t tiles, cudaStream_t stream, cudaStream_t* long_stream) override {
    *long_stream = nullptr;
    cudaError_t e = cudaSuccess;
    if (!ls) {
      int lo = 0, hi = 0;
      (void)cudaDeviceGetStreamPriorityRange(&lo, &hi);
      e = cudaStreamCreateWithPriority(&ls, cudaStreamNonBlocking, hi);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&fork, cudaEventDisableTiming);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&join, cudaEventDisableTiming);
      if (e != cudaSuccess) return e;
    }
    if (tiles != n_tiles) {
      release();
      n_tiles = tiles, cap = std::min<uint32_t>(std::max<uint32_t>(tiles / 16u, 64u), 4096u);
      for (int k = 0; k < 2 && e == cudaSuccess; ++k) {  // one allocation per set: [flags (n_tiles) | count | list (cap)] -- flags and count are zeroed together
        e = cudaMalloc(&flag[k], ((size_t)n_tiles + 1 + cap) * 4);
        list[k] = flag[k] ? flag[k] + n_tiles : nullptr;
      }
      if (e != cudaSuccess) {
        release();
        return e;
      }
    }
    const int nxt = cur ^ 1;
    e = cudaMemsetAsync(flag[nxt], 0, ((size_t)n_tiles + 1) * 4, stream);  // the flags and the count behind them
    if (e != cudaSuccess) return e;
    P.next_list = list[nxt], P.next_flag = flag[nxt];
    P.sched_cap = cap;  // (P.sched_threshold: the launcher's option)
    prepared = true;
    P.prev_list = nullptr, P.prev_flag = nullptr;
    if (valid) {
      P.prev_list = list[cur], P.prev_flag = flag[cur];
      e = cudaEventRecord(fork, stream);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(ls, fork, 0);
      if (e != cudaSuccess) return e;
      *long_stream = ls;
    }
    return cudaSuccess;
  }
  void attach(RenderParams& P) override {
    P.prev_list = nullptr, P.prev_flag = nullptr, P.next_list = nullptr, P.next_flag = nullptr;
    if (!prepared) return;
    P.next_list = list[cur ^ 1], P.next_flag = flag[cur ^ 1], P.sched_cap = cap;
    if (valid) P.prev_list = list[cur], P.prev_flag = flag[cur];
  }
  cudaError_t finish(cudaStream_t stream) override {
    cudaError_t e = cudaEventRecord(join, ls);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(stream, join, 0);
    return e;
  }
  void launched() { cur ^= 1, valid = true, prepared = false; }  // the set just recorded becomes "previous"
  bool prepared = false;  // prepare() has run for the frame being launched
};
constexpr size_t kSchedEntries = 24;

struct DeviceSlot {
  int id = 0;
  std::vector<std::shared_ptr<SchedEntry>> sched;  // long-tiles-first state per (geometry, stream)
  uint64_t sched_clock = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool events_pending = false;
  bool peer_to_first = false;  // can store straight into device[0] memory
  WxState* d_states = nullptr;
  uint32_t states_cap = 0;
  uint8_t* scratch = nullptr;  // staging frame when peer stores are impossible
  size_t scratch_bytes = 0;
  uint32_t resident_ctas = 0;    // SM count x CTAs per SM
  cudaStream_t copy_stream = nullptr;   // read-back of finished row chunks while later chunks render
  cudaStream_t aux[2] = {nullptr, nullptr};  // chunks alternate over stream/aux[0]/aux[1] so that one chunk's drain overlaps the next
  cudaEvent_t fork = nullptr, join[2] = {nullptr, nullptr};
  std::vector<cudaEvent_t> chunk_done;  // one per chunk of a pipelined wx_render
};

struct FrameBuffers {  // device[0]-resident outputs of wx_render
  uint8_t* rgba = nullptr;
  size_t rgba_bytes = 0;
  size_t rgba_valid = 0;  // bytes of the last rendered frame(s)
  // The last multi-device wx_render left every device's row bands in that device's own frame (read back over each
  // device's own PCIe link); device 0's frame is completed over NVLink when something needs it whole (gather_frame).
  bool distributed = false;
  uint32_t dist_states = 0, dist_width = 0, dist_height = 0;
  uint8_t* rgb = nullptr;  // wx_capture_srgb staging
  size_t rgb_bytes = 0;
  void* aov[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t aov_bytes[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

struct WxContext {
  std::vector<DeviceSlot> dev;
  FrameBuffers fb;
  cudaEvent_t total0 = nullptr, total1 = nullptr;
  bool total_pending = false;
  WxRenderInfo info{};
  std::string last_error;
  LaunchOptions opt;           // wx_set_option
  uint32_t render_chunks = 0;  // WX_OPT_RENDER_CHUNKS (0 = automatic)
  bool nvtx = false;           // WX_OPT_NVTX
};

struct TreeOnDevice {
  uint32_t *e5 = nullptr, *e4 = nullptr;
  uint8_t* l3 = nullptr;
  int4* origins = nullptr;
  uint32_t *grid = nullptr, *f4 = nullptr;  // world grid and re-encoded N4 tables (wx_device.cuh), derived from e5 / e4 on the device
};

struct WxTree {
  WxContext* ctx = nullptr;
  std::vector<TreeOnDevice> on;  // one per context device
  std::vector<int4> origins;  // biased by kBias (wx_device.cuh)
  int16_t root_grid[64];      // root cells of [-8192, 8192)^3 (DevTree::root_grid)
  uint32_t leaf_shift = 9;    // log2(bytes per leaf brick)
  bool fast_ok = true;        // every step size < 2^20: the fast march applies
  bool grid_ok = true;        // ... and the index bases of the world grid fit 32 bits: march_grid applies
  int32_t bbox_cells[6] = {1 << 30, 1 << 30, 1 << 30, -1, -1, -1};  // grid_bbox_cells (tolerance-mode march)
  WxTreeInfo info{};
};

// NVTX range around a phase of an entry point (upload / sweep / render / read-back), only when WX_OPT_NVTX is set: the
// reference's tracing hooks are its wgpu debug labels; these show up as named ranges in Nsight Systems.
struct NvtxRange {
  bool on;
  NvtxRange(const WxContext* ctx, const char* name) : on(ctx && ctx->nvtx) {
    if (on) nvtxRangePushA(name);
  }
  ~NvtxRange() {
    if (on) nvtxRangePop();
  }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

static thread_local std::string g_last_error;  // failures before a context exists

static int fail(WxContext* ctx, int status, const std::string& what) {
  if (ctx) ctx->last_error = what;
  g_last_error = what;
  return status;
}
static int fail_cuda(WxContext* ctx, cudaError_t e, const char* where) {
  std::string msg = std::string(where) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
  (void)cudaGetLastError();  // clear the sticky-free error state
  return fail(ctx, e == cudaErrorMemoryAllocation ? WX_ERR_OUT_OF_MEMORY : WX_ERR_CUDA, msg);
}
#define WX_CUDA(ctx, call)                                      \
  do {                                                          \
    cudaError_t e_ = (call);                                    \
    if (e_ != cudaSuccess) return fail_cuda((ctx), e_, #call);  \
  } while (0)

// ---------------------------------------------------------------------------------------------
// World grid + re-encoded N4 tables of one replica, derived on the device from the e5 / e4 tables it already holds (uploaded
// by wx_tree_upload or written by the SDF sweep of wx_tree_build).  Same formulas as the host version in wx_pack.h
// (build_grid_tables), through the helpers of wx_device.cuh.
// ---------------------------------------------------------------------------------------------
namespace {
struct RootGrid {
  int16_t v[64];
};
__global__ void n4_origins_kernel(const uint32_t* __restrict__ e5, const int4* __restrict__ origins, uint32_t n5, uint32_t* __restrict__ o4) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n5 * 32768u) return;
  const uint32_t e = e5[i];
  if (!(e & kChildFlag)) return;
  const int4 o = origins[i >> 15];
  const uint32_t s = (uint32_t)i & 32767u;
  uint32_t* d = o4 + (size_t)(e & ~kChildFlag) * 3u;
  d[0] = (uint32_t)o.x + (s >> 10) * 128u, d[1] = (uint32_t)o.y + ((s >> 5) & 31u) * 128u, d[2] = (uint32_t)o.z + (s & 31u) * 128u;
}
__global__ void build_grid_kernel(const uint32_t* __restrict__ e5, const __grid_constant__ RootGrid rg, uint32_t* __restrict__ grid) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;  // the 128^3 cells of the cube (the pads keep the memset's kEntrySlow)
  if (c >= kGS * kGS2) return;
  grid[(size_t)kGridPad + c] = grid_cell_entry(c >> 14, (c >> 7) & 127u, c & 127u, rg.v, e5);
}
// bbox[0..2] = min, bbox[3..5] = max cell coordinate of the in-world cells that are children or active tiles (grid_bbox_cells)
__global__ void grid_bbox_kernel(const uint32_t* __restrict__ grid, int32_t* __restrict__ bbox) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= 64u * 64u * 64u) return;
  const uint32_t cx = 32u + (c >> 12), cy = 32u + ((c >> 6) & 63u), cz = 32u + (c & 63u);
  const uint32_t e = grid[(size_t)kGridPad + (size_t)cx * kGS2 + (size_t)cy * kGS + cz];
  if (!((e & kChildFlag) || e == 0u)) return;
  atomicMin(bbox + 0, (int32_t)cx), atomicMin(bbox + 1, (int32_t)cy), atomicMin(bbox + 2, (int32_t)cz);
  atomicMax(bbox + 3, (int32_t)cx), atomicMax(bbox + 4, (int32_t)cy), atomicMax(bbox + 5, (int32_t)cz);
}
__global__ void build_f4_kernel(const uint32_t* __restrict__ e4, const uint32_t* __restrict__ o4, uint32_t n4, uint32_t* __restrict__ f4) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n4 * 4096u) return;
  const uint32_t e = e4[i];
  const uint32_t* o = o4 + (i >> 12) * 3u;
  const uint32_t s = (uint32_t)i & 4095u;
  f4[i] = (e & kChildFlag) ? grid_word3(e & ~kChildFlag, o[0] + (s >> 8) * 8u, o[1] + ((s >> 4) & 15u) * 8u, o[2] + (s & 15u) * 8u) : e;
}
}  // namespace

// Current device = the replica's device; everything is enqueued on `stream` (the scratch is freed in stream order).
// bbox_host (optional, 6 ints): receives grid_bbox_cells of the new grid (the copy is stream-ordered into pageable memory, i.e.
// complete when this returns).
static cudaError_t build_grid_on_device(uint32_t*& grid, uint32_t*& f4, const uint32_t* e5, const uint32_t* e4, const int4* origins,
                                        uint32_t n5, uint32_t n4, const int16_t root_grid[64], cudaStream_t stream,
                                        int32_t* bbox_host = nullptr) {
  cudaError_t e = cudaMalloc(&grid, kGridCells * 4);
  if (e == cudaSuccess) e = cudaMalloc(&f4, (size_t)n4 * 4096u * 4u + 256);
  uint32_t* o4 = nullptr;
  if (e == cudaSuccess) e = cudaMallocAsync((void**)&o4, (size_t)n4 * 12u + 256, stream);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(grid, 0x02, kGridCells * 4, stream);  // kEntrySlow everywhere
  if (e == cudaSuccess && n5) {
    n4_origins_kernel<<<(unsigned)(((size_t)n5 * 32768u + 255) / 256), 256, 0, stream>>>(e5, origins, n5, o4);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) {
    RootGrid rg;
    memcpy(rg.v, root_grid, sizeof(rg.v));
    build_grid_kernel<<<kGS * kGS2 / 256, 256, 0, stream>>>(e5, rg, grid);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess && n4) {
    build_f4_kernel<<<(unsigned)(((size_t)n4 * 4096u + 255) / 256), 256, 0, stream>>>(e4, o4, n4, f4);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess && bbox_host) {
    int32_t* bb = reinterpret_cast<int32_t*>(o4);  // the scratch is free again once build_f4_kernel has run (stream order)
    const int32_t init[6] = {1 << 30, 1 << 30, 1 << 30, -1, -1, -1};
    e = cudaMemcpyAsync(bb, init, sizeof(init), cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) {
      grid_bbox_kernel<<<64 * 64 * 64 / 256, 256, 0, stream>>>(grid, bb);
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(bbox_host, bb, sizeof(init), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  }
  const cudaError_t fe = cudaFreeAsync(o4, stream);
  return e != cudaSuccess ? e : fe;
}

extern "C" int wx_abi_version(void) { return WX_ABI_VERSION; }

extern "C" const char* wx_strerror(int status) {
  switch (status) {
    case WX_OK: return "ok";
    case WX_ERR_INVALID_ARGUMENT: return "invalid argument";
    case WX_ERR_NO_DEVICE: return "no CUDA device available (this path has no CPU fallback)";
    case WX_ERR_CUDA: return "CUDA runtime error";
    case WX_ERR_OUT_OF_MEMORY: return "out of device memory";
    case WX_ERR_BAD_TREE: return "inconsistent tree description";
    case WX_ERR_UNSUPPORTED: return "unsupported input";
    case WX_ERR_PEER_ACCESS: return "peer access unavailable";
    default: return "unknown status";
  }
}

extern "C" const char* wx_last_error(const WxContext* ctx) { return ctx ? ctx->last_error.c_str() : g_last_error.c_str(); }

extern "C" int wx_init(int n_devices, const int* device_ids, WxContext** out) try {
  if (!out || n_devices < 0) return fail(nullptr, WX_ERR_INVALID_ARGUMENT, "wx_init: bad arguments");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    (void)cudaGetLastError();
    return fail(nullptr, WX_ERR_NO_DEVICE, std::string("wx_init: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "0 devices"));
  }
  WxContext* ctx = new (std::nothrow) WxContext();
  if (!ctx) return WX_ERR_OUT_OF_MEMORY;
  int current = 0;
  (void)cudaGetDevice(&current);
  const int n = n_devices == 0 ? 1 : n_devices;
  for (int i = 0; i < n; ++i) {
    DeviceSlot s;
    s.id = n_devices == 0 ? current : (device_ids ? device_ids[i] : i);
    if (s.id < 0 || s.id >= count) {
      delete ctx;
      return fail(nullptr, WX_ERR_INVALID_ARGUMENT, "wx_init: device id out of range");
    }
    ctx->dev.push_back(s);
  }
  for (size_t i = 0; i < ctx->dev.size(); ++i) {
    DeviceSlot& s = ctx->dev[i];
    cudaError_t err = cudaSetDevice(s.id);
    if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&s.copy_stream, cudaStreamNonBlocking);
    if (err == cudaSuccess) {
      int sms = 0;
      err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s.id);
      s.resident_ctas = (uint32_t)sms * kCtasPerSm;
    }
    for (int k = 0; k < 2 && err == cudaSuccess; ++k) {
      err = cudaStreamCreateWithFlags(&s.aux[k], cudaStreamNonBlocking);
      if (err == cudaSuccess) err = cudaEventCreateWithFlags(&s.join[k], cudaEventDisableTiming);
    }
    if (err == cudaSuccess) err = cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming);
    if (err == cudaSuccess) err = cudaEventCreate(&s.ev0);
    if (err == cudaSuccess) err = cudaEventCreate(&s.ev1);
    if (err != cudaSuccess) {
      int rc = fail_cuda(nullptr, err, "wx_init: stream/event creation");
      wx_shutdown(ctx);
      return rc;
    }
    if (i > 0 && s.id == ctx->dev[0].id) {
      s.peer_to_first = true;  // the same GPU listed again (a second set of streams, frame and tree replica): plain stores
    } else if (i > 0) {
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, s.id, ctx->dev[0].id) == cudaSuccess && can) {
        cudaError_t pe = cudaDeviceEnablePeerAccess(ctx->dev[0].id, 0);
        s.peer_to_first = (pe == cudaSuccess || pe == cudaErrorPeerAccessAlreadyEnabled);
      }
      (void)cudaGetLastError();
    }
  }
  (void)cudaSetDevice(ctx->dev[0].id);
  if (cudaEventCreate(&ctx->total0) != cudaSuccess || cudaEventCreate(&ctx->total1) != cudaSuccess) {
    int rc = fail_cuda(nullptr, cudaGetLastError(), "wx_init: event creation");
    wx_shutdown(ctx);
    return rc;
  }
  *out = ctx;
  return WX_OK;
} catch (const std::bad_alloc&) {
  return fail(nullptr, WX_ERR_OUT_OF_MEMORY, "wx_init: host allocation failed");
} catch (const std::exception& ex) {
  return fail(nullptr, WX_ERR_UNSUPPORTED, std::string("wx_init: ") + ex.what());
} catch (...) {
  return fail(nullptr, WX_ERR_UNSUPPORTED, "wx_init: unexpected exception");
}

static void free_frame_buffers(WxContext* ctx) {
  if (ctx->dev.empty()) return;
  (void)cudaSetDevice(ctx->dev[0].id);
  if (ctx->fb.rgba) (void)cudaFree(ctx->fb.rgba);
  if (ctx->fb.rgb) (void)cudaFree(ctx->fb.rgb);
  for (auto& p : ctx->fb.aov)
    if (p) (void)cudaFree(p);
  ctx->fb = FrameBuffers();
}

extern "C" int wx_shutdown(WxContext* ctx) {
  if (!ctx) return WX_OK;
  free_frame_buffers(ctx);
  for (DeviceSlot& s : ctx->dev) {
    (void)cudaSetDevice(s.id);
    if (s.stream) (void)cudaStreamSynchronize(s.stream);
    (void)cudaDeviceSynchronize();
    s.sched.clear();
    if (s.d_states) (void)cudaFree(s.d_states);
    if (s.scratch) (void)cudaFree(s.scratch);
    if (s.ev0) (void)cudaEventDestroy(s.ev0);
    if (s.ev1) (void)cudaEventDestroy(s.ev1);
    for (cudaEvent_t e : s.chunk_done) (void)cudaEventDestroy(e);
    if (s.copy_stream) (void)cudaStreamDestroy(s.copy_stream);
    for (int k = 0; k < 2; ++k) {
      if (s.aux[k]) (void)cudaStreamDestroy(s.aux[k]);
      if (s.join[k]) (void)cudaEventDestroy(s.join[k]);
    }
    if (s.fork) (void)cudaEventDestroy(s.fork);
    if (s.stream) (void)cudaStreamDestroy(s.stream);
  }
  if (ctx->total0) (void)cudaEventDestroy(ctx->total0);
  if (ctx->total1) (void)cudaEventDestroy(ctx->total1);
  delete ctx;
  return WX_OK;
}

extern "C" int wx_device_count(const WxContext* ctx) { return ctx ? (int)ctx->dev.size() : 0; }

// ---------------------------------------------------------------------------------------------
// Tree packing: WxTreeDesc (reference order: masks + per-slot u32) -> entry tables + leaf bricks
// ---------------------------------------------------------------------------------------------
// Host-side part of a device tree: biased origins, root cells, limits, info.
static WxTree* new_tree(WxContext* ctx, const WxTreeDesc* d, uint32_t leaf_bits, uint32_t max5, uint32_t max4, uint32_t max3v) {
  WxTree* t = new (std::nothrow) WxTree();
  if (!t) return nullptr;
  t->ctx = ctx;
  bias_origins(d->n5, d->origins, t->origins);
  build_root_grid(d->n5, d->origins, t->root_grid);
  t->leaf_shift = leaf_bits == 8 ? 9 : 11;
  // the fast march needs byte leaves and every step size below 2^20 (wx_device.cuh); anything else takes the exact march
  t->fast_ok = fast_march_ok(leaf_bits, max5, max4, max3v);
  t->grid_ok = world_grid_ok(t->fast_ok, d->n4, d->n3);
  t->info.n5 = d->n5, t->info.n4 = d->n4, t->info.n3 = d->n3;
  t->info.leaf_bits = leaf_bits;
  t->info.max_dist[0] = max5, t->info.max_dist[1] = max4, t->info.max_dist[2] = max3v;
  t->info.n_devices = (uint32_t)ctx->dev.size();
  t->info.device_bytes = ((size_t)d->n5 * 32768 + (size_t)d->n4 * 4096) * 4 + ((size_t)d->n3 << t->leaf_shift) + t->origins.size() * sizeof(int4) +
                         (t->grid_ok ? kGridCells * 4 + (size_t)d->n4 * 4096 * 4 : 0);
  t->on.resize(ctx->dev.size());
  return t;
}

extern "C" int wx_tree_upload(WxContext* ctx, const WxTreeDesc* d, WxTree** out) try {
  if (!ctx || !d || !out) return fail(ctx, WX_ERR_INVALID_ARGUMENT, "wx_tree_upload: null argument");
  *out = nullptr;
  if ((d->n5 && (!d->origins || !d->kids5 || !d->vals5 || !d->tab5)) || (d->n4 && (!d->kids4 || !d->vals4 || !d->tab4)) ||
      (d->n3 && (!d->vals3 || !d->tab3)))
    return fail(ctx, WX_ERR_INVALID_ARGUMENT, "wx_tree_upload: missing array");
  if (d->n3 && d->tab3_elem_bytes != 1 && d->tab3_elem_bytes != 4)
    return fail(ctx, WX_ERR_INVALID_ARGUMENT, "wx_tree_upload: tab3_elem_bytes must be 1 or 4");
  if (d->n5 > (uint32_t)kRootIndexMask || d->n4 >= kChildFlag || d->n3 >= kChildFlag)
    return fail(ctx, WX_ERR_UNSUPPORTED, "wx_tree_upload: too many nodes");  // 16383 N5s would be 2 GB of N5 tables alone
  NvtxRange range_all(ctx, "wx_tree_upload");
  std::unique_ptr<NvtxRange> range_pack(new (std::nothrow) NvtxRange(ctx, "wx_tree_upload: pack (host)"));

  std::vector<uint32_t> e5, e4;
  std::vector<uint8_t> l3;
  uint32_t max5 = 0, max4 = 0;
  int rc = pack_internal(d->n5, 32768, 128.f, d->kids5, d->vals5, d->tab5, d->n4, e5, &max5);
  if (rc) return fail(ctx, rc, "wx_tree_upload: N5 table (child index out of range or distance >= 2^31)");
  rc = pack_internal(d->n4, 4096, 8.f, d->kids4, d->vals4, d->tab4, d->n3, e4, &max4);
  if (rc) return fail(ctx, rc, "wx_tree_upload: N4 table (child index out of range or distance >= 2^31)");

  // leaf distance width: the largest inactive-voxel distance decides (one byte per voxel, else u32)
  uint32_t max3v = 0;
  const uint32_t leaf_bits = pack_leaves(d->n3, d->vals3, d->tab3, d->tab3_elem_bytes, l3, &max3v);

  WxTree* t = new_tree(ctx, d, leaf_bits, max5, max4, max3v);
  range_pack.reset();
  NvtxRange range_up(ctx, "wx_tree_upload: H2D + world grid");
  if (!t) return fail(ctx, WX_ERR_OUT_OF_MEMORY, "wx_tree_upload: host allocation");

  auto up = [&](int dev_i) -> int {
    DeviceSlot& s = ctx->dev[dev_i];
    TreeOnDevice& o = t->on[dev_i];
    WX_CUDA(ctx, cudaSetDevice(s.id));
    // +256 B of slack keeps zero-sized levels allocatable
    WX_CUDA(ctx, cudaMalloc(&o.e5, e5.size() * 4 + 256));
    WX_CUDA(ctx, cudaMalloc(&o.e4, e4.size() * 4 + 256));
    WX_CUDA(ctx, cudaMalloc(&o.l3, l3.size() + 256));
    WX_CUDA(ctx, cudaMalloc(&o.origins, t->origins.size() * sizeof(int4) + 256));
    if (!e5.empty()) WX_CUDA(ctx, cudaMemcpyAsync(o.e5, e5.data(), e5.size() * 4, cudaMemcpyHostToDevice, s.stream));
    if (!e4.empty()) WX_CUDA(ctx, cudaMemcpyAsync(o.e4, e4.data(), e4.size() * 4, cudaMemcpyHostToDevice, s.stream));
    if (!l3.empty()) WX_CUDA(ctx, cudaMemcpyAsync(o.l3, l3.data(), l3.size(), cudaMemcpyHostToDevice, s.stream));
    if (!t->origins.empty())
      WX_CUDA(ctx, cudaMemcpyAsync(o.origins, t->origins.data(), t->origins.size() * sizeof(int4), cudaMemcpyHostToDevice, s.stream));
    if (t->grid_ok) WX_CUDA(ctx, build_grid_on_device(o.grid, o.f4, o.e5, o.e4, o.origins, d->n5, d->n4, t->root_grid, s.stream, dev_i == 0 ? t->bbox_cells : nullptr));
    return WX_OK;
  };
  for (int i = 0; i < (int)ctx->dev.size(); ++i) {
    rc = up(i);
    if (rc) {
      wx_tree_free(ctx, t);
      return rc;
    }
  }
  for (DeviceSlot& s : ctx->dev) {
    (void)cudaSetDevice(s.id);
    cudaError_t e = cudaStreamSynchronize(s.stream);
    if (e != cudaSuccess) {
      wx_tree_free(ctx, t);
      return fail_cuda(ctx, e, "wx_tree_upload: synchronize");
    }
  }
  (void)cudaSetDevice(ctx->dev[0].id);
  *out = t;
  return WX_OK;
} catch (const std::bad_alloc&) {
  return fail(nullptr, WX_ERR_OUT_OF_MEMORY, "wx_tree_upload: host allocation failed");
} catch (const std::exception& ex) {
  return fail(nullptr, WX_ERR_UNSUPPORTED, std::string("wx_tree_upload: ") + ex.what());
} catch (...) {
  return fail(nullptr, WX_ERR_UNSUPPORTED, "wx_tree_upload: unexpected exception");
}

static int check_topology(WxContext* ctx, const WxTreeDesc* d, const char* who) {
  for (size_t i = 0; i < (size_t)d->n5 * 32768; ++i)
    if (bit(d->kids5, i) && d->tab5[i] >= d->n4) return fail(ctx, WX_ERR_BAD_TREE, std::string(who) + ": N5 child index out of range");
  for (size_t i = 0; i < (size_t)d->n4 * 4096; ++i)
    if (bit(d->kids4, i) && d->tab4[i] >= d->n3) return fail(ctx, WX_ERR_BAD_TREE, std::string(who) + ": N4 child index out of range");
  return WX_OK;
}

extern "C" int wx_compute_sdf(WxContext* ctx, const WxTreeDesc* d, uint32_t* tab5_out, uint32_t* tab4_out, void* tab3_out,
                              uint32_t tab3_elem_bytes, WxSdfInfo* info) try {
  if (!ctx || !d) return fail(ctx, WX_ERR_INVALID_ARGUMENT, "wx_compute_sdf: null argument");
  if ((d->n5 && (!d->origins || !d->kids5 || !d->tab5 || !tab5_out)) || (d->n4 && (!d->kids4 || !d->tab4 || !tab4_out)) ||
      (d->n3 && (!d->vals3 || !tab3_out)))
    return fail(ctx, WX_ERR_INVALID_ARGUMENT, "wx_compute_sdf: missing array");
  if (tab3_elem_bytes != 1 && tab3_elem_bytes != 4) return fail(ctx, WX_ERR_INVALID_ARGUMENT, "wx_compute_sdf: tab3_elem_bytes must be 1 or 4");
  // child indices must be in range: the sweeps follow them
  if (int bad = check_topology(ctx, d, "wx_compute_sdf")) return bad;
  NvtxRange range(ctx, "wx_compute_sdf: sweep");
  DeviceSlot& d0 = ctx->dev[0];
  WX_CUDA(ctx, cudaSetDevice(d0.id));
  const auto t0 = std::chrono::steady_clock::now();
  uint32_t r[5] = {0, 0, 0, 0, 0};
  float ms = 0.f;
  WX_CUDA(ctx, compute_sdf_device(*d, tab5_out, tab4_out, tab3_out, tab3_elem_bytes, r, &ms, d0.stream));
  if (info) {
    info->max_dist[0] = r[0], info->max_dist[1] = r[1], info->max_dist[2] = r[2], info->rounds = r[4];
    info->device_ms = ms;
    info->total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
  }
  if (r[3]) return fail(ctx, WX_ERR_UNSUPPORTED, "wx_compute_sdf: a leaf distance does not fit the requested element size");
  return WX_OK;
} catch (const std::bad_alloc&) {
  return fail(nullptr, WX_ERR_OUT_OF_MEMORY, "wx_compute_sdf: host allocation failed");
} catch (const std::exception& ex) {
  return fail(nullptr, WX_ERR_UNSUPPORTED, std::string("wx_compute_sdf: ") + ex.what());
} catch (...) {
  return fail(nullptr, WX_ERR_UNSUPPORTED, "wx_compute_sdf: unexpected exception");
}

extern "C" int wx_tree_build(WxContext* ctx, const WxTreeDesc* d, WxTree** out, WxSdfInfo* info) try {
  if (!ctx || !d || !out) return fail(ctx, WX_ERR_INVALID_ARGUMENT, "wx_tree_build: null argument");
  *out = nullptr;
  if ((d->n5 && (!d->origins || !d->kids5 || !d->vals5 || !d->tab5)) || (d->n4 && (!d->kids4 || !d->vals4 || !d->tab4)) || (d->n3 && !d->vals3))
    return fail(ctx, WX_ERR_INVALID_ARGUMENT, "wx_tree_build: missing array");
  if (d->n5 > (uint32_t)kRootIndexMask || d->n4 >= kChildFlag || d->n3 >= kChildFlag) return fail(ctx, WX_ERR_UNSUPPORTED, "wx_tree_build: too many nodes");
  int rc = check_topology(ctx, d, "wx_tree_build");
  if (rc) return rc;
  NvtxRange range(ctx, "wx_tree_build: sweep + pack + replicate");
  const auto t0 = std::chrono::steady_clock::now();
  const size_t s5 = (size_t)d->n5 * 32768, s4 = (size_t)d->n4 * 4096, s3 = (size_t)d->n3 * 512;
  DeviceSlot& d0 = ctx->dev[0];
  WX_CUDA(ctx, cudaSetDevice(d0.id));
  TreeOnDevice first;
  auto drop = [&](TreeOnDevice& o) {
    if (o.e5) (void)cudaFree(o.e5);
    if (o.e4) (void)cudaFree(o.e4);
    if (o.l3) (void)cudaFree(o.l3);
    if (o.origins) (void)cudaFree(o.origins);
    o = TreeOnDevice();
  };
  cudaError_t e = cudaMalloc(&first.e5, s5 * 4 + 256);
  if (e == cudaSuccess) e = cudaMalloc(&first.e4, s4 * 4 + 256);
  if (e == cudaSuccess) e = cudaMalloc(&first.l3, s3 + 256);
  if (e == cudaSuccess) e = cudaMalloc(&first.origins, (size_t)d->n5 * sizeof(int4) + 256);
  uint32_t r[5] = {0, 0, 0, 0, 0};
  float ms = 0.f;
  if (e == cudaSuccess) {
    const SdfDeviceTargets targets{first.e5, first.e4, first.l3};
    e = compute_sdf_device(*d, nullptr, nullptr, nullptr, 1, r, &ms, d0.stream, &targets);
  }
  if (e != cudaSuccess) {
    drop(first);
    return fail_cuda(ctx, e, "wx_tree_build");
  }
  if (info) {
    info->max_dist[0] = r[0], info->max_dist[1] = r[1], info->max_dist[2] = r[2], info->rounds = r[4];
    info->device_ms = ms;
  }
  if (r[3]) {  // a leaf distance above 255: this tree needs the u32 brick layout
    drop(first);
    return fail(ctx, WX_ERR_UNSUPPORTED, "wx_tree_build: a leaf distance exceeds 255 (use wx_compute_sdf + wx_tree_upload)");
  }
  WxTree* t = new_tree(ctx, d, 8, r[0], r[1], r[2]);
  if (!t) {
    drop(first);
    return fail(ctx, WX_ERR_OUT_OF_MEMORY, "wx_tree_build: host allocation");
  }
  t->on[0] = first;
  auto bail = [&](cudaError_t err, const char* where) {
    wx_tree_free(ctx, t);
    return fail_cuda(ctx, err, where);
  };
  if (!t->origins.empty()) {
    e = cudaMemcpyAsync(first.origins, t->origins.data(), t->origins.size() * sizeof(int4), cudaMemcpyHostToDevice, d0.stream);
    if (e != cudaSuccess) return bail(e, "wx_tree_build: origins");
  }
  if (t->grid_ok) {
    e = build_grid_on_device(t->on[0].grid, t->on[0].f4, first.e5, first.e4, first.origins, d->n5, d->n4, t->root_grid, d0.stream, t->bbox_cells);
    if (e != cudaSuccess) return bail(e, "wx_tree_build: world grid");
  }
  // replicate read-only on the other devices of the context (peer copies over NVLink)
  for (size_t i = 1; i < ctx->dev.size(); ++i) {
    DeviceSlot& s = ctx->dev[i];
    TreeOnDevice& o = t->on[i];
    e = cudaSetDevice(s.id);
    if (e == cudaSuccess) e = cudaMalloc(&o.e5, s5 * 4 + 256);
    if (e == cudaSuccess) e = cudaMalloc(&o.e4, s4 * 4 + 256);
    if (e == cudaSuccess) e = cudaMalloc(&o.l3, s3 + 256);
    if (e == cudaSuccess) e = cudaMalloc(&o.origins, t->origins.size() * sizeof(int4) + 256);
    if (e == cudaSuccess && t->grid_ok) e = cudaMalloc(&o.grid, kGridCells * 4);
    if (e == cudaSuccess && t->grid_ok) e = cudaMalloc(&o.f4, s4 * 4 + 256);
    if (e == cudaSuccess) e = cudaSetDevice(d0.id);
    if (e == cudaSuccess && s5) e = cudaMemcpyPeerAsync(o.e5, s.id, first.e5, d0.id, s5 * 4, d0.stream);
    if (e == cudaSuccess && s4) e = cudaMemcpyPeerAsync(o.e4, s.id, first.e4, d0.id, s4 * 4, d0.stream);
    if (e == cudaSuccess && s3) e = cudaMemcpyPeerAsync(o.l3, s.id, first.l3, d0.id, s3, d0.stream);
    if (e == cudaSuccess && !t->origins.empty())
      e = cudaMemcpyPeerAsync(o.origins, s.id, first.origins, d0.id, t->origins.size() * sizeof(int4), d0.stream);
    if (e == cudaSuccess && t->grid_ok) {
      e = cudaMemcpyPeerAsync(o.grid, s.id, t->on[0].grid, d0.id, kGridCells * 4, d0.stream);
      if (e == cudaSuccess && s4) e = cudaMemcpyPeerAsync(o.f4, s.id, t->on[0].f4, d0.id, s4 * 4, d0.stream);
    }
    if (e != cudaSuccess) return bail(e, "wx_tree_build: replicate");
  }
  e = cudaStreamSynchronize(d0.stream);
  if (e != cudaSuccess) return bail(e, "wx_tree_build: synchronize");
  if (info) info->total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
  *out = t;
  return WX_OK;
} catch (const std::bad_alloc&) {
  return fail(nullptr, WX_ERR_OUT_OF_MEMORY, "wx_tree_build: host allocation failed");
} catch (const std::exception& ex) {
  return fail(nullptr, WX_ERR_UNSUPPORTED, std::string("wx_tree_build: ") + ex.what());
} catch (...) {
  return fail(nullptr, WX_ERR_UNSUPPORTED, "wx_tree_build: unexpected exception");
}

extern "C" int wx_tree_free(WxContext* ctx, WxTree* tree) {
  if (!tree) return WX_OK;
  WxContext* c = ctx ? ctx : tree->ctx;
  for (size_t i = 0; i < tree->on.size() && c && i < c->dev.size(); ++i) {
    (void)cudaSetDevice(c->dev[i].id);
    (void)cudaDeviceSynchronize();  // launches of wx_render_device (any stream) may still be reading the tables
    TreeOnDevice& o = tree->on[i];
    if (o.e5) (void)cudaFree(o.e5);
    if (o.e4) (void)cudaFree(o.e4);
    if (o.l3) (void)cudaFree(o.l3);
    if (o.origins) (void)cudaFree(o.origins);
    if (o.grid) (void)cudaFree(o.grid);
    if (o.f4) (void)cudaFree(o.f4);
  }
  delete tree;
  return WX_OK;
}

extern "C" int wx_tree_info(const WxTree* tree, WxTreeInfo* info) {
  if (!tree || !info) return WX_ERR_INVALID_ARGUMENT;
  *info = tree->info;
  return WX_OK;
}

// ---------------------------------------------------------------------------------------------
// Frame entry points
// ---------------------------------------------------------------------------------------------
// Frames [cam0, cam0 + ncam) of `states`, rows [row0, row1) (row1 == 0: all rows; with a shard row0 must be a multiple of
// one round of the band deal, shard->count * shard->band_rows).
// `states_on_device`: the whole batch is already in s.d_states (a pipelined wx_render uploads it once).
static SchedEntry* find_sched(DeviceSlot& s, const SchedKey& key);
static int launch_on(WxContext* ctx, int dev_i, const WxTree* tree, const WxState* states, uint32_t n_states, uint32_t width,
                     uint32_t height, uint8_t* rgba_dev, const WxAov* aov_dev, const WxShard* shard, cudaStream_t stream,
                     uint32_t* launches_out, uint32_t cam0 = 0, uint32_t ncam = 0xffffffffu, uint32_t row0 = 0, uint32_t row1 = 0,
                     bool states_on_device = false, bool long_first = true, SchedEntry* frame_sched = nullptr,
                     const SchedCall* sched_call = nullptr) {
  DeviceSlot& s = ctx->dev[dev_i];
  const TreeOnDevice& o = tree->on[dev_i];
  WxState* launch_states = nullptr;
  RenderParams P;
  memset(&P, 0, sizeof(P));
  fill_dev_tree(P.tree, o.e5, o.e4, o.l3, o.origins, tree->info.n5, tree->info.n4, tree->info.n3, tree->leaf_shift, tree->fast_ok,
                tree->root_grid, o.grid, o.f4, ctx->opt.march == 2 ? tree->bbox_cells : nullptr);  // the clip only with WX_OPT_MARCH = 2
  P.n_states = n_states;
  P.width = width, P.height = height;
  P.rgba = reinterpret_cast<uchar4*>(rgba_dev);
  if (aov_dev) {
    P.aov.state = aov_dev->state, P.aov.voxel = aov_dev->voxel, P.aov.leaf = aov_dev->leaf, P.aov.level = aov_dev->level;
    P.aov.iters = aov_dev->iters, P.aov.depth = aov_dev->depth, P.aov.mask = aov_dev->mask, P.aov.pos = aov_dev->pos;
    P.has_aov = (P.aov.state || P.aov.voxel || P.aov.leaf || P.aov.level || P.aov.iters || P.aov.depth || P.aov.mask || P.aov.pos) ? 1u : 0u;
  }
  if (n_states == 1) {
    P.s0 = states[0];
  } else {
    if (!states_on_device) {
      // The batch of THIS launch, allocated and freed in stream order: asynchronous wx_render_device calls on different
      // streams never share a states buffer (s.d_states is only used by the blocking wx_render, which uploads it once).
      WX_CUDA(ctx, cudaMallocAsync((void**)&launch_states, (size_t)n_states * sizeof(WxState), stream));
      cudaError_t ce = cudaMemcpyAsync(launch_states, states, (size_t)n_states * sizeof(WxState), cudaMemcpyHostToDevice, stream);
      if (ce != cudaSuccess) {
        (void)cudaFreeAsync(launch_states, stream);
        return fail_cuda(ctx, ce, "render: states upload");
      }
      P.states = launch_states;
    } else {
      P.states = s.d_states;
    }
  }
  const uint32_t cam1 = ncam > n_states - cam0 ? n_states : cam0 + ncam;
  uint32_t total_launches = 0;
  // a camera batch is launched per run of equal render modes (normally one run)
  for (uint32_t b = cam0; b < cam1;) {
    uint32_t e = b + 1;
    const uint32_t mode = states[b].render_mode[0] > 4 ? 0 : states[b].render_mode[0];
    while (e < cam1 && (states[e].render_mode[0] > 4 ? 0 : states[e].render_mode[0]) == mode) ++e;
    P.cam_base = b;
    if (shard) P.shard_index = shard->index, P.shard_count = shard->count, P.band_rows = shard->band_rows;
    else P.shard_index = 0, P.shard_count = 1, P.band_rows = 0;
    P.row_base = row0, P.row_end = row1;
    uint32_t l = 0;
    // work-queue head of a persistent kernel (WX_OPT_KERNEL != 0): allocated and freed in stream order, so that launches on
    // different streams never share one however many are in flight (a ring of counters could be reused too early)
    uint32_t* counter = nullptr;
    if (ctx->opt.kernel != 0) {
      const cudaError_t ae = cudaMallocAsync((void**)&counter, sizeof(uint32_t), stream);
      if (ae != cudaSuccess) {
        if (launch_states) (void)cudaFreeAsync(launch_states, stream);
        return fail_cuda(ctx, ae, "render: work counter");
      }
    }
    // long-tiles-first state of this launch geometry on this stream (plain launches when the option is off)
    SchedEntry* se = frame_sched;
    P.prev_list = nullptr, P.prev_flag = nullptr, P.next_list = nullptr, P.next_flag = nullptr;
    if (!se && long_first && ctx->opt.long_first && ctx->opt.kernel == 0) {
      SchedKey key;
      memset(&key, 0, sizeof(key));
      key.tree = tree, key.stream = stream, key.width = width, key.height = height, key.cam0 = b, key.ncam = e - b;
      key.shard_index = P.shard_index, key.shard_count = P.shard_count, key.band_rows = P.band_rows, key.row0 = row0, key.row1 = row1;
      se = find_sched(s, key);
    }
    const cudaError_t le = launch_raycast(P, e - b, mode, stream, &l, counter, s.resident_ctas, ctx->opt, se, sched_call);
    if (counter) (void)cudaFreeAsync(counter, stream);  // behind the kernel that used it
    if (le == cudaSuccess && se && P.next_list && !sched_call) se->launched();  // (a chunked frame: the caller, after its last chunk)
    if (le != cudaSuccess) {
      if (launch_states) (void)cudaFreeAsync(launch_states, stream);
      return fail_cuda(ctx, le, "launch_raycast");
    }
    total_launches += l;
    b = e;
  }
  if (launch_states) WX_CUDA(ctx, cudaFreeAsync(launch_states, stream));
  *launches_out = total_launches;
  return WX_OK;
}

// The long-tiles-first entry of a launch geometry on a stream (created on first use; the entry used longest ago makes room).
static SchedEntry* find_sched(DeviceSlot& s, const SchedKey& key) {
  SchedEntry* se = nullptr;
  for (auto& c : s.sched)
    if (c->key == key) se = c.get();
  if (!se) {
    if (s.sched.size() >= kSchedEntries) {
      size_t victim = 0;
      for (size_t k = 1; k < s.sched.size(); ++k)
        if (s.sched[k]->last_use < s.sched[victim]->last_use) victim = k;
      s.sched.erase(s.sched.begin() + (long)victim);
    }
    s.sched.emplace_back(new SchedEntry());
    se = s.sched.back().get();
    se->key = key;
  }
  se->last_use = ++s.sched_clock;
  return se;
}

static bool shard_ok(const WxShard* shard) {
  return !shard || (shard->count != 0 && shard->index < shard->count &&
                    (shard->count == 1 || (shard->band_rows != 0 && shard->band_rows % kBandRowsMultiple == 0)));
}

static int check_render_args(WxContext* ctx, const WxTree* tree, const WxState* states, uint32_t n_states, uint32_t width,
                             uint32_t height, const void* rgba) {
  if (!ctx || !tree || !states || !rgba) return fail(ctx, WX_ERR_INVALID_ARGUMENT, "render: null argument");
  if (tree->ctx != ctx) return fail(ctx, WX_ERR_INVALID_ARGUMENT, "render: tree belongs to another context");
  if (n_states == 0 || width == 0 || height == 0) return fail(ctx, WX_ERR_INVALID_ARGUMENT, "render: empty frame");
  if ((uint64_t)width * height * n_states > (1ull << 40)) return fail(ctx, WX_ERR_INVALID_ARGUMENT, "render: frame too large");
  return WX_OK;
}

extern "C" int wx_render_device(WxContext* ctx, int device_index, const WxTree* tree, const WxState* states, uint32_t n_states,
                                uint32_t width, uint32_t height, uint8_t* rgba_dev, const WxAov* aov_dev, const WxShard* shard,
                                void* stream) try {
  int rc = check_render_args(ctx, tree, states, n_states, width, height, rgba_dev);
  if (rc) return rc;
  if (device_index < 0 || device_index >= (int)ctx->dev.size()) return fail(ctx, WX_ERR_INVALID_ARGUMENT, "render: device index");
  if (!shard_ok(shard))
    return fail(ctx, WX_ERR_INVALID_ARGUMENT, "render: bad shard (band_rows must be a positive multiple of 8)");
  NvtxRange range(ctx, "wx_render_device: launch");
  DeviceSlot& s = ctx->dev[device_index];
  WX_CUDA(ctx, cudaSetDevice(s.id));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  WX_CUDA(ctx, cudaEventRecord(s.ev0, st));
  uint32_t launches = 0;
  rc = launch_on(ctx, device_index, tree, states, n_states, width, height, rgba_dev, aov_dev, shard, st, &launches);
  if (rc) return rc;
  WX_CUDA(ctx, cudaEventRecord(s.ev1, st));
  for (DeviceSlot& d : ctx->dev) d.events_pending = false;
  s.events_pending = true;
  ctx->total_pending = false;
  ctx->info = WxRenderInfo{};
  ctx->info.launches = launches;
  ctx->info.rays = (uint64_t)(width / 8 * 8) * (height / 4 * 4) * n_states;  // whole frame; a shard renders its share
  return WX_OK;
} catch (const std::bad_alloc&) {
  return fail(nullptr, WX_ERR_OUT_OF_MEMORY, "wx_render_device: host allocation failed");
} catch (const std::exception& ex) {
  return fail(nullptr, WX_ERR_UNSUPPORTED, std::string("wx_render_device: ") + ex.what());
} catch (...) {
  return fail(nullptr, WX_ERR_UNSUPPORTED, "wx_render_device: unexpected exception");
}

// Pageable host memory: a device-to-host cudaMemcpyAsync into it returns only when the copy is done (the driver stages it), so a
// call that alternates launches and copies would serialise them.  The pipelined paths then enqueue every kernel first and copy
// afterwards; chunks still leave in order while later ones render, but overlap of the copies themselves needs wx_host_alloc_pinned.
static bool host_pageable(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    (void)cudaGetLastError();
    return true;
  }
  return at.type == cudaMemoryTypeUnregistered;
}

static int ensure(WxContext* ctx, void** p, size_t* have, size_t need) {
  if (*have >= need) return WX_OK;
  if (*p) (void)cudaFree(*p);
  *p = nullptr, *have = 0;
  WX_CUDA(ctx, cudaMalloc(p, need));
  *have = need;
  return WX_OK;
}


// The bands of device `i` in rows [row0, row1) (row0 a multiple of one round of the deal) of frames [cam0, cam1): one
// strided copy per frame (pitch = one round of the deal) plus the last, partial band when it falls to this device.
// `dst` and `src` have the frame layout.
static cudaError_t copy_own_bands(uint8_t* dst, const uint8_t* src, int i, int ndev, uint32_t cam0, uint32_t cam1, uint32_t row0,
                                  uint32_t row1, uint32_t width, uint32_t height, cudaMemcpyKind kind, cudaStream_t st) {
  const size_t band_bytes = (size_t)kBandRowsMultiple * width * 4, frame_bytes = (size_t)height * width * 4;
  const uint32_t rows = row1 - row0, full = rows / kBandRowsMultiple, tail_rows = rows % kBandRowsMultiple;
  const uint32_t own_full = full > (uint32_t)i ? (full - (uint32_t)i + (uint32_t)ndev - 1) / (uint32_t)ndev : 0;
  for (uint32_t c = cam0; c < cam1; ++c) {
    const size_t base = (size_t)c * frame_bytes + (size_t)row0 * width * 4, off = base + (size_t)i * band_bytes;
    if (own_full) {
      cudaError_t e = cudaMemcpy2DAsync(dst + off, (size_t)ndev * band_bytes, src + off, (size_t)ndev * band_bytes, band_bytes, own_full, kind, st);
      if (e != cudaSuccess) return e;
    }
    if (tail_rows && full % (uint32_t)ndev == (uint32_t)i) {
      const size_t toff = base + (size_t)full * band_bytes;
      cudaError_t e = cudaMemcpyAsync(dst + toff, src + toff, (size_t)tail_rows * width * 4, kind, st);
      if (e != cudaSuccess) return e;
    }
  }
  return cudaSuccess;
}

// wx_render on several devices without AOVs.  Device i renders its bands into its own frame (device 0: the context's
// frame), camera group by camera group over three kernel streams, and each finished group leaves for the host on the
// device's copy stream.  Returns with every copy complete and total1 recorded on device 0's stream.
static int render_distributed(WxContext* ctx, const WxTree* tree, const WxState* states, uint32_t n_states, uint32_t width,
                              uint32_t height, uint8_t* rgba_out, uint32_t* launches) {
  const int ndev = (int)ctx->dev.size();
  const size_t npix = (size_t)n_states * width * height;
  DeviceSlot& d0 = ctx->dev[0];
  // Pipeline units: groups of cameras, or -- one large frame -- 4 / 2 / 2 row blocks on 2 / 4 / 8 devices, each whole
  // rounds of the deal (more blocks cost more in host calls per device than the overlap returns: profiles/r1_multi_device_n8.txt).
  struct Chunk {
    uint32_t cam0, cam1, row0, row1;
  };
  std::vector<Chunk> chunks;
  if (n_states > 1) {
    const uint32_t k = std::min<uint32_t>(n_states, 16u), per = (n_states + k - 1) / k;
    for (uint32_t c = 0; c < n_states; c += per) chunks.push_back(Chunk{c, std::min(n_states, c + per), 0, height});
  } else {
    const uint32_t k = ctx->render_chunks ? ctx->render_chunks : ((uint64_t)width * height >= (1u << 20) ? std::max(2u, 8u / (uint32_t)ndev) : 1u);
    const uint32_t round = (uint32_t)ndev * kBandRowsMultiple, rows = ((height + k - 1) / k + round - 1) / round * round;
    for (uint32_t r = 0; r < height; r += rows) chunks.push_back(Chunk{0, 1, r, std::min(height, r + rows)});
  }
  const uint32_t n_chunks = (uint32_t)chunks.size();
  // Pageable destination: pass 0 enqueues every device's kernels, pass 1 copies (host_pageable); else one pass does both.
  const bool deferred = host_pageable(rgba_out);
  for (int pass = 0; pass < (deferred ? 2 : 1); ++pass)
  for (int i = 0; i < ndev; ++i) {
    DeviceSlot& s = ctx->dev[i];
    WX_CUDA(ctx, cudaSetDevice(s.id));
    uint8_t* local = ctx->fb.rgba;
    if (i != 0) {
      int rc = ensure(ctx, (void**)&s.scratch, &s.scratch_bytes, npix * 4);
      if (rc) return rc;
      local = s.scratch;
    }
    if (pass == 1) {  // the copies of a pageable destination, each behind its chunk's kernel
      if (n_chunks == 1) {
        WX_CUDA(ctx, copy_own_bands(rgba_out, local, i, ndev, 0, n_states, 0, height, width, height, cudaMemcpyDefault, s.stream));
      } else {
        for (uint32_t c_idx = 0; c_idx < n_chunks; ++c_idx) {
          const Chunk& ch = chunks[c_idx];
          WX_CUDA(ctx, cudaStreamWaitEvent(s.copy_stream, s.chunk_done[c_idx], 0));
          WX_CUDA(ctx, copy_own_bands(rgba_out, local, i, ndev, ch.cam0, ch.cam1, ch.row0, ch.row1, width, height, cudaMemcpyDefault, s.copy_stream));
        }
        WX_CUDA(ctx, cudaEventRecord(s.fork, s.copy_stream));
        WX_CUDA(ctx, cudaStreamWaitEvent(s.stream, s.fork, 0));
      }
      WX_CUDA(ctx, cudaEventRecord(s.join[0], s.stream));  // kernels and copies of this device
      continue;
    }
    while (s.chunk_done.size() < n_chunks) {
      cudaEvent_t e;
      WX_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      s.chunk_done.push_back(e);
    }
    if (i != 0) WX_CUDA(ctx, cudaStreamWaitEvent(s.stream, ctx->total0, 0));  // the call starts when device 0's stream gets here
    if (n_states > 1) {
      if (s.states_cap < n_states) {
        if (s.d_states) (void)cudaFree(s.d_states);
        s.d_states = nullptr, s.states_cap = 0;
        WX_CUDA(ctx, cudaMalloc(&s.d_states, (size_t)n_states * sizeof(WxState)));
        s.states_cap = n_states;
      }
      WX_CUDA(ctx, cudaMemcpyAsync(s.d_states, states, (size_t)n_states * sizeof(WxState), cudaMemcpyHostToDevice, s.stream));
    }
    WX_CUDA(ctx, cudaEventRecord(s.ev0, s.stream));
    WxShard sh{(uint32_t)i, (uint32_t)ndev, (uint32_t)kBandRowsMultiple, 0};
    s.events_pending = true;
    if (n_chunks == 1) {  // kernel and copy in stream order, nothing to overlap (and the fewest host calls)
      uint32_t l = 0;
      int rc = launch_on(ctx, i, tree, states, n_states, width, height, local, nullptr, &sh, s.stream, &l, 0, 0xffffffffu, 0, 0, n_states > 1);
      if (rc) return rc;
      *launches += l;
      WX_CUDA(ctx, cudaEventRecord(s.ev1, s.stream));
      if (!deferred) WX_CUDA(ctx, copy_own_bands(rgba_out, local, i, ndev, 0, n_states, 0, height, width, height, cudaMemcpyDefault, s.stream));
    } else {
      WX_CUDA(ctx, cudaEventRecord(s.fork, s.stream));
      cudaStream_t ks[3] = {s.stream, s.aux[0], s.aux[1]};
      for (int k = 0; k < 2; ++k) WX_CUDA(ctx, cudaStreamWaitEvent(s.aux[k], s.fork, 0));
      for (uint32_t c_idx = 0; c_idx < n_chunks; ++c_idx) {
        const Chunk& ch = chunks[c_idx];
        cudaStream_t st = ks[c_idx % 3];
        uint32_t l = 0;
        int rc = launch_on(ctx, i, tree, states, n_states, width, height, local, nullptr, &sh, st, &l, ch.cam0, ch.cam1 - ch.cam0, ch.row0,
                           ch.row1, n_states > 1, false);
        if (rc) return rc;
        *launches += l;
        WX_CUDA(ctx, cudaEventRecord(s.chunk_done[c_idx], st));
        if (!deferred) WX_CUDA(ctx, cudaStreamWaitEvent(s.copy_stream, s.chunk_done[c_idx], 0));
        if (!deferred) WX_CUDA(ctx, copy_own_bands(rgba_out, local, i, ndev, ch.cam0, ch.cam1, ch.row0, ch.row1, width, height, cudaMemcpyDefault, s.copy_stream));
      }
      for (int k = 0; k < 2; ++k) {
        WX_CUDA(ctx, cudaEventRecord(s.join[k], s.aux[k]));
        WX_CUDA(ctx, cudaStreamWaitEvent(s.stream, s.join[k], 0));
      }
      WX_CUDA(ctx, cudaEventRecord(s.ev1, s.stream));  // all kernels of this device
      if (!deferred) {
        WX_CUDA(ctx, cudaEventRecord(s.fork, s.copy_stream));
        WX_CUDA(ctx, cudaStreamWaitEvent(s.stream, s.fork, 0));
      }
    }
    if (!deferred) WX_CUDA(ctx, cudaEventRecord(s.join[0], s.stream));  // kernels and copies of this device
  }
  WX_CUDA(ctx, cudaSetDevice(d0.id));
  for (int i = 1; i < ndev; ++i) WX_CUDA(ctx, cudaStreamWaitEvent(d0.stream, ctx->dev[i].join[0], 0));
  return WX_OK;
}

// Completes device 0's frame after a distributed render: every other device's bands travel over NVLink (peer copy).
static int gather_frame(WxContext* ctx) {
  if (!ctx->fb.distributed) return WX_OK;
  const int ndev = (int)ctx->dev.size();
  DeviceSlot& d0 = ctx->dev[0];
  for (int i = 1; i < ndev; ++i) {
    DeviceSlot& s = ctx->dev[i];
    WX_CUDA(ctx, cudaSetDevice(s.id));
    WX_CUDA(ctx, copy_own_bands(ctx->fb.rgba, s.scratch, i, ndev, 0, ctx->fb.dist_states, 0, ctx->fb.dist_height, ctx->fb.dist_width,
                                ctx->fb.dist_height, cudaMemcpyDefault, s.stream));
    WX_CUDA(ctx, cudaEventRecord(s.join[0], s.stream));
  }
  WX_CUDA(ctx, cudaSetDevice(d0.id));
  for (int i = 1; i < ndev; ++i) WX_CUDA(ctx, cudaStreamWaitEvent(d0.stream, ctx->dev[i].join[0], 0));
  ctx->fb.distributed = false;
  return WX_OK;
}

extern "C" int wx_render(WxContext* ctx, const WxTree* tree, const WxState* states, uint32_t n_states, uint32_t width,
                         uint32_t height, uint8_t* rgba_out, const WxAov* aov_out) try {
  int rc = check_render_args(ctx, tree, states, n_states, width, height, rgba_out);
  if (rc) return rc;
  const size_t npix = (size_t)n_states * width * height;
  DeviceSlot& d0 = ctx->dev[0];
  WX_CUDA(ctx, cudaSetDevice(d0.id));
  rc = ensure(ctx, (void**)&ctx->fb.rgba, &ctx->fb.rgba_bytes, npix * 4);
  if (rc) return rc;
  // AOV staging on device 0
  NvtxRange range_all(ctx, "wx_render");
  std::unique_ptr<NvtxRange> range_launch(new (std::nothrow) NvtxRange(ctx, "wx_render: launches (+ pipelined read-back)"));
  static const size_t aov_elem[8] = {1, 12, 4, 1, 4, 4, 1, 12};
  void* host_aov[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  WxAov dev_aov;
  memset(&dev_aov, 0, sizeof(dev_aov));
  bool any_aov = false;
  if (aov_out) {
    host_aov[0] = aov_out->state, host_aov[1] = aov_out->voxel, host_aov[2] = aov_out->leaf, host_aov[3] = aov_out->level;
    host_aov[4] = aov_out->iters, host_aov[5] = aov_out->depth, host_aov[6] = aov_out->mask, host_aov[7] = aov_out->pos;
    for (int k = 0; k < 8; ++k) {
      if (!host_aov[k]) continue;
      any_aov = true;
      rc = ensure(ctx, &ctx->fb.aov[k], &ctx->fb.aov_bytes[k], npix * aov_elem[k]);
      if (rc) return rc;
    }
    dev_aov.state = host_aov[0] ? (uint8_t*)ctx->fb.aov[0] : nullptr;
    dev_aov.voxel = host_aov[1] ? (int32_t*)ctx->fb.aov[1] : nullptr;
    dev_aov.leaf = host_aov[2] ? (int32_t*)ctx->fb.aov[2] : nullptr;
    dev_aov.level = host_aov[3] ? (uint8_t*)ctx->fb.aov[3] : nullptr;
    dev_aov.iters = host_aov[4] ? (uint32_t*)ctx->fb.aov[4] : nullptr;
    dev_aov.depth = host_aov[5] ? (float*)ctx->fb.aov[5] : nullptr;
    dev_aov.mask = host_aov[6] ? (uint8_t*)ctx->fb.aov[6] : nullptr;
    dev_aov.pos = host_aov[7] ? (float*)ctx->fb.aov[7] : nullptr;
  }

  const int ndev = (int)ctx->dev.size();
  WX_CUDA(ctx, cudaEventRecord(ctx->total0, d0.stream));
  uint32_t launches = 0;
  bool copied = false;  // the frame has already been read back chunk by chunk
  if (ndev == 1) {
    WX_CUDA(ctx, cudaEventRecord(d0.ev0, d0.stream));
    // Pipelined read-back: the frame is rendered in chunks (row blocks of one frame, or whole frames of a camera
    // batch); chunk k travels to the host on the copy stream while chunk k+1 renders.
    struct Chunk {
      uint32_t cam0, ncam, row0, row1;
    };
    std::vector<Chunk> chunks;
    if (!any_aov) {
      if (n_states == 1) {
        const uint32_t k = ctx->render_chunks ? ctx->render_chunks : ((uint64_t)width * height >= (1u << 20) ? 8u : 1u);
        const uint32_t rows = ((height + k - 1) / k + kBandRowsMultiple - 1) / kBandRowsMultiple * kBandRowsMultiple;
        for (uint32_t r = 0; r < height; r += rows) chunks.push_back(Chunk{0, 1, r, std::min(height, r + rows)});
      } else {
        const uint32_t per = (n_states + 63) / 64;
        for (uint32_t c = 0; c < n_states; c += per) chunks.push_back(Chunk{c, std::min(per, n_states - c), 0, height});
      }
    }
    if (chunks.size() <= 1) {
      rc = launch_on(ctx, 0, tree, states, n_states, width, height, ctx->fb.rgba, any_aov ? &dev_aov : nullptr, nullptr, d0.stream, &launches);
      if (rc) return rc;
    } else {
      while (d0.chunk_done.size() < chunks.size()) {
        cudaEvent_t e;
        WX_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        d0.chunk_done.push_back(e);
      }
      if (n_states > 1) {  // upload the batch once; the chunk launches read it in place
        if (d0.states_cap < n_states) {
          if (d0.d_states) (void)cudaFree(d0.d_states);
          d0.d_states = nullptr, d0.states_cap = 0;
          WX_CUDA(ctx, cudaMalloc(&d0.d_states, (size_t)n_states * sizeof(WxState)));
          d0.states_cap = n_states;
        }
        WX_CUDA(ctx, cudaMemcpyAsync(d0.d_states, states, (size_t)n_states * sizeof(WxState), cudaMemcpyHostToDevice, d0.stream));
      }
      // A chunk ends with a drain as long as its longest ray; chunks therefore alternate over three streams, so
      // that the next chunk fills the SMs the previous one is leaving.
      // Long tiles first, once per FRAME: a list per chunk launch costs a second launch and a fork / join each (measured 0.855 vs
      // 0.820 ms of kernels per 4K frame), so the whole frame's long tiles are rendered by one long-tile kernel ahead of the chunk
      // launches, which skip them (tile ids are the frame's) and record the next frame's list between them.
      SchedEntry* fse = nullptr;
      bool long_launched = false;
      SchedCall call_chunk;
      call_chunk.phase = kSchedChunk, call_chunk.frame_tile_rows = (height + raycast_tile_height() - 1) / raycast_tile_height();
      if (n_states == 1 && ctx->opt.long_first && ctx->opt.kernel == 0 && raycast_tile_height() == (uint32_t)kBandRowsMultiple) {
        SchedKey key;
        memset(&key, 0, sizeof(key));
        key.tree = tree, key.stream = d0.stream, key.width = width, key.height = height, key.cam0 = 0, key.ncam = 1;
        key.shard_index = 0, key.shard_count = 1;
        fse = find_sched(d0, key);
        SchedCall call_long;
        call_long.phase = kSchedLongOnly;
        uint32_t l = 0;
        rc = launch_on(ctx, 0, tree, states, n_states, width, height, ctx->fb.rgba, nullptr, nullptr, d0.stream, &l, 0, 0xffffffffu, 0, 0, true, true,
                       fse, &call_long);
        if (rc) return rc;
        launches += l;
        long_launched = l != 0;
        if (!fse->prepared) fse = nullptr;  // not applicable to this frame (render mode, size): plain chunk launches
        else if (long_launched) WX_CUDA(ctx, fse->finish(d0.copy_stream));  // no chunk leaves before the long tiles are rendered
      }
      WX_CUDA(ctx, cudaEventRecord(d0.fork, d0.stream));
      cudaStream_t ks[3] = {d0.stream, d0.aux[0], d0.aux[1]};
      for (int k = 0; k < 2; ++k) WX_CUDA(ctx, cudaStreamWaitEvent(d0.aux[k], d0.fork, 0));
      const bool deferred = host_pageable(rgba_out);
      for (size_t c = 0; c < chunks.size(); ++c) {
        const Chunk& ch = chunks[c];
        cudaStream_t st = ks[c % 3];
        uint32_t l = 0;
        call_chunk.tile_row_offset = ch.row0 / raycast_tile_height();
        rc = launch_on(ctx, 0, tree, states, n_states, width, height, ctx->fb.rgba, nullptr, nullptr, st, &l, ch.cam0, ch.ncam,
                       ch.row0, ch.row1, true, false, fse, fse ? &call_chunk : nullptr);
        if (rc) return rc;
        launches += l;
        WX_CUDA(ctx, cudaEventRecord(d0.chunk_done[c], st));
        if (deferred) continue;  // pageable destination: every launch first (host_pageable)
        WX_CUDA(ctx, cudaStreamWaitEvent(d0.copy_stream, d0.chunk_done[c], 0));
        const size_t off = ((size_t)ch.cam0 * height + ch.row0) * width * 4;
        const size_t bytes = ch.ncam > 1 || ch.row1 - ch.row0 == height ? (size_t)ch.ncam * height * width * 4 : (size_t)(ch.row1 - ch.row0) * width * 4;
        WX_CUDA(ctx, cudaMemcpyAsync(rgba_out + off, ctx->fb.rgba + off, bytes, cudaMemcpyDefault, d0.copy_stream));  // host, or any GPU's memory (UVA)
      }
      for (size_t c = 0; deferred && c < chunks.size(); ++c) {
        const Chunk& ch = chunks[c];
        WX_CUDA(ctx, cudaStreamWaitEvent(d0.copy_stream, d0.chunk_done[c], 0));
        const size_t off = ((size_t)ch.cam0 * height + ch.row0) * width * 4;
        const size_t bytes = ch.ncam > 1 || ch.row1 - ch.row0 == height ? (size_t)ch.ncam * height * width * 4 : (size_t)(ch.row1 - ch.row0) * width * 4;
        WX_CUDA(ctx, cudaMemcpyAsync(rgba_out + off, ctx->fb.rgba + off, bytes, cudaMemcpyDefault, d0.copy_stream));
      }
      // join: the main stream waits for the other kernel streams (ev1 = all kernels done), then for the last copy
      for (int k = 0; k < 2; ++k) {
        WX_CUDA(ctx, cudaEventRecord(d0.join[k], d0.aux[k]));
        WX_CUDA(ctx, cudaStreamWaitEvent(d0.stream, d0.join[k], 0));
      }
      if (fse) {
        if (long_launched) WX_CUDA(ctx, fse->finish(d0.stream));
        fse->launched();
      }
      WX_CUDA(ctx, cudaEventRecord(d0.ev1, d0.stream));
      WX_CUDA(ctx, cudaEventRecord(d0.fork, d0.copy_stream));
      WX_CUDA(ctx, cudaStreamWaitEvent(d0.stream, d0.fork, 0));
      copied = true;
    }
    if (!copied) WX_CUDA(ctx, cudaEventRecord(d0.ev1, d0.stream));
    d0.events_pending = true;
  } else if (!any_aov) {
    // Row bands of one tile height, dealt round-robin: every GPU gets hit-heavy and empty regions.  Each GPU renders its
    // bands into its own frame and sends them to the host over its own PCIe link while its next cameras render; the
    // eight links together deliver a frame ~3x faster than device 0's alone.  Device 0's frame is completed lazily.
    rc = render_distributed(ctx, tree, states, n_states, width, height, rgba_out, &launches);
    if (rc) return rc;
    copied = true;
  } else {
    // AOVs requested: every GPU stores its pixels straight into device 0's frame and AOV planes over NVLink (the store
    // is the gather), and device 0 reads everything back.
    for (int i = 0; i < ndev; ++i) {
      DeviceSlot& s = ctx->dev[i];
      WX_CUDA(ctx, cudaSetDevice(s.id));
      WxShard sh{(uint32_t)i, (uint32_t)ndev, (uint32_t)kBandRowsMultiple, 0};
      uint8_t* dst = ctx->fb.rgba;
      const bool direct = (i == 0) || s.peer_to_first;
      if (!direct) {
        if (any_aov) return fail(ctx, WX_ERR_PEER_ACCESS, "wx_render: AOVs on several devices need peer access to device 0");
        rc = ensure(ctx, (void**)&s.scratch, &s.scratch_bytes, npix * 4);
        if (rc) return rc;
        dst = s.scratch;
      }
      WX_CUDA(ctx, cudaEventRecord(s.ev0, s.stream));
      uint32_t l = 0;
      rc = launch_on(ctx, i, tree, states, n_states, width, height, dst, any_aov ? &dev_aov : nullptr, &sh, s.stream, &l);
      if (rc) return rc;
      launches += l;
      WX_CUDA(ctx, cudaEventRecord(s.ev1, s.stream));
      s.events_pending = true;
      if (!direct) {  // staged gather: copy this GPU's bands into device 0's frame
        const size_t band_bytes = (size_t)kBandRowsMultiple * width * 4;
        const uint32_t bands = (height + kBandRowsMultiple - 1) / kBandRowsMultiple;
        for (uint32_t c = 0; c < n_states; ++c)
          for (uint32_t b = (uint32_t)i; b < bands; b += (uint32_t)ndev) {
            const size_t off = ((size_t)c * height + (size_t)b * kBandRowsMultiple) * width * 4;
            const size_t bytes = std::min(band_bytes, ((size_t)(c + 1) * height * width * 4) - off);
            WX_CUDA(ctx, cudaMemcpyPeerAsync(ctx->fb.rgba + off, d0.id, s.scratch + off, s.id, bytes, s.stream));
          }
      }
    }
    for (int i = 1; i < ndev; ++i) {  // device 0's stream waits for the peers' stores
      WX_CUDA(ctx, cudaSetDevice(ctx->dev[i].id));
      WX_CUDA(ctx, cudaEventRecord(ctx->dev[i].join[0], ctx->dev[i].stream));  // the slot's own event (created in wx_init)
      WX_CUDA(ctx, cudaSetDevice(d0.id));
      WX_CUDA(ctx, cudaStreamWaitEvent(d0.stream, ctx->dev[i].join[0], 0));
    }
    WX_CUDA(ctx, cudaSetDevice(d0.id));
  }
  if (!copied) WX_CUDA(ctx, cudaMemcpyAsync(rgba_out, ctx->fb.rgba, npix * 4, cudaMemcpyDefault, d0.stream));
  range_launch.reset();
  NvtxRange range_rb(ctx, "wx_render: read-back + synchronize");
  for (int k = 0; k < 8; ++k)
    if (host_aov[k]) WX_CUDA(ctx, cudaMemcpyAsync(host_aov[k], ctx->fb.aov[k], npix * aov_elem[k], cudaMemcpyDeviceToHost, d0.stream));
  WX_CUDA(ctx, cudaEventRecord(ctx->total1, d0.stream));
  WX_CUDA(ctx, cudaStreamSynchronize(d0.stream));
  ctx->total_pending = true;
  ctx->fb.rgba_valid = npix * 4;
  ctx->fb.distributed = ndev > 1 && !any_aov;
  ctx->fb.dist_states = n_states, ctx->fb.dist_width = width, ctx->fb.dist_height = height;
  ctx->info = WxRenderInfo{};
  ctx->info.launches = launches;
  ctx->info.rays = (uint64_t)(width / 8 * 8) * (height / 4 * 4) * n_states;
  return WX_OK;
} catch (const std::bad_alloc&) {
  return fail(nullptr, WX_ERR_OUT_OF_MEMORY, "wx_render: host allocation failed");
} catch (const std::exception& ex) {
  return fail(nullptr, WX_ERR_UNSUPPORTED, std::string("wx_render: ") + ex.what());
} catch (...) {
  return fail(nullptr, WX_ERR_UNSUPPORTED, "wx_render: unexpected exception");
}

extern "C" int wx_render_shard(WxContext* ctx, const WxTree* tree, const WxState* states, uint32_t n_states, uint32_t width,
                               uint32_t height, const WxShard* shard, uint8_t* rgba_out) try {
  int rc = check_render_args(ctx, tree, states, n_states, width, height, rgba_out);
  if (rc) return rc;
  if (!shard || !shard_ok(shard) || shard->band_rows != (uint32_t)kBandRowsMultiple)
    return fail(ctx, WX_ERR_INVALID_ARGUMENT, "wx_render_shard: a shard with band_rows == 8 is required");
  NvtxRange range(ctx, "wx_render_shard: render + deliver rows");
  const size_t npix = (size_t)n_states * width * height;
  DeviceSlot& d0 = ctx->dev[0];
  WX_CUDA(ctx, cudaSetDevice(d0.id));
  rc = ensure(ctx, (void**)&ctx->fb.rgba, &ctx->fb.rgba_bytes, npix * 4);
  if (rc) return rc;
  WX_CUDA(ctx, cudaEventRecord(ctx->total0, d0.stream));
  WX_CUDA(ctx, cudaEventRecord(d0.ev0, d0.stream));
  uint32_t launches = 0;
  rc = launch_on(ctx, 0, tree, states, n_states, width, height, ctx->fb.rgba, nullptr, shard, d0.stream, &launches);
  if (rc) return rc;
  WX_CUDA(ctx, cudaEventRecord(d0.ev1, d0.stream));
  WX_CUDA(ctx, copy_own_bands(rgba_out, ctx->fb.rgba, (int)shard->index, (int)shard->count, 0, n_states, 0, height, width, height,
                              cudaMemcpyDefault, d0.stream));
  WX_CUDA(ctx, cudaEventRecord(ctx->total1, d0.stream));
  WX_CUDA(ctx, cudaStreamSynchronize(d0.stream));
  for (DeviceSlot& d : ctx->dev) d.events_pending = false;
  d0.events_pending = true;
  ctx->total_pending = true;
  ctx->fb.rgba_valid = 0;  // only a shard's rows are valid: nothing for wx_capture_srgb
  ctx->fb.distributed = false;
  ctx->info = WxRenderInfo{};
  ctx->info.launches = launches;
  ctx->info.rays = (uint64_t)(width / 8 * 8) * (height / 4 * 4) * n_states;
  return WX_OK;
} catch (const std::bad_alloc&) {
  return fail(nullptr, WX_ERR_OUT_OF_MEMORY, "wx_render_shard: host allocation failed");
} catch (const std::exception& ex) {
  return fail(nullptr, WX_ERR_UNSUPPORTED, std::string("wx_render_shard: ") + ex.what());
} catch (...) {
  return fail(nullptr, WX_ERR_UNSUPPORTED, "wx_render_shard: unexpected exception");
}

extern "C" int wx_set_option(WxContext* ctx, int option, int64_t value) {
  if (!ctx) return fail(nullptr, WX_ERR_INVALID_ARGUMENT, "wx_set_option: null context");
  switch (option) {
    case WX_OPT_MARCH:
      if (value < 0 || value > 2) break;
      ctx->opt.march = (int)value;
      return WX_OK;
    case WX_OPT_KERNEL:
      if (value < 0 || value > 2) break;
      ctx->opt.kernel = (int)value;
      return WX_OK;
    case WX_OPT_RENDER_CHUNKS:
      if (value < 0 || value > 4096) break;
      ctx->render_chunks = (uint32_t)value;
      return WX_OK;
    case WX_OPT_SMEM_PAD:
      if (value < 0 || value > 200 * 1024) break;
      ctx->opt.smem_pad = (size_t)value;
      return WX_OK;
    case WX_OPT_NVTX:
      if (value != 0 && value != 1) break;
      ctx->nvtx = value != 0;
      return WX_OK;
    case WX_OPT_LONG_FIRST:
      if (value != 0 && value != 1) break;
      ctx->opt.long_first = (int)value;
      return WX_OK;
    case WX_OPT_LONG_THRESHOLD:
      if (value < 1 || value > 1000) break;
      ctx->opt.long_threshold = (uint32_t)value;
      return WX_OK;
    default:
      return fail(ctx, WX_ERR_INVALID_ARGUMENT, "wx_set_option: unknown option");
  }
  return fail(ctx, WX_ERR_INVALID_ARGUMENT, "wx_set_option: value out of range");
}

extern "C" int wx_get_option(const WxContext* ctx, int option, int64_t* value_out) {
  if (!ctx || !value_out) return WX_ERR_INVALID_ARGUMENT;
  switch (option) {
    case WX_OPT_MARCH: *value_out = ctx->opt.march; return WX_OK;
    case WX_OPT_KERNEL: *value_out = ctx->opt.kernel; return WX_OK;
    case WX_OPT_RENDER_CHUNKS: *value_out = ctx->render_chunks; return WX_OK;
    case WX_OPT_SMEM_PAD: *value_out = (int64_t)ctx->opt.smem_pad; return WX_OK;
    case WX_OPT_NVTX: *value_out = ctx->nvtx ? 1 : 0; return WX_OK;
    case WX_OPT_LONG_FIRST: *value_out = ctx->opt.long_first; return WX_OK;
    case WX_OPT_LONG_THRESHOLD: *value_out = ctx->opt.long_threshold; return WX_OK;
    default: return WX_ERR_INVALID_ARGUMENT;
  }
}

extern "C" int wx_shard_rows(uint32_t height, const WxShard* shard, uint8_t* row_mask_out) {
  if (!row_mask_out || !shard_ok(shard)) return fail(nullptr, WX_ERR_INVALID_ARGUMENT, "wx_shard_rows: bad shard (band_rows must be a positive multiple of 8)");
  memset(row_mask_out, 0, height);
  if (!shard || shard->count == 1) {
    memset(row_mask_out, 1, height);
    return WX_OK;
  }
  const uint32_t own = shard_own_bands(height, shard->index, shard->count, shard->band_rows);
  for (uint32_t k = 0; k < own; ++k) {
    const uint64_t first = (uint64_t)(k * shard->count + shard->index) * shard->band_rows;
    for (uint64_t y = first; y < first + shard->band_rows && y < height; ++y) row_mask_out[y] = 1;
  }
  return WX_OK;
}

extern "C" int wx_srgb_table(uint8_t table_out[256]) {
  if (!table_out) return WX_ERR_INVALID_ARGUMENT;
  build_srgb_lut(table_out);
  return WX_OK;
}

extern "C" int wx_capture_srgb(WxContext* ctx, uint32_t n_states, uint32_t width, uint32_t height, uint8_t* rgb_out) try {
  if (!ctx || !rgb_out) return fail(ctx, WX_ERR_INVALID_ARGUMENT, "wx_capture_srgb: null argument");
  const size_t npix = (size_t)n_states * width * height;
  if (npix == 0 || npix * 4 != ctx->fb.rgba_valid) return fail(ctx, WX_ERR_INVALID_ARGUMENT, "wx_capture_srgb: no frame of that size was rendered last");
  NvtxRange range(ctx, "wx_capture_srgb");
  DeviceSlot& d0 = ctx->dev[0];
  WX_CUDA(ctx, cudaSetDevice(d0.id));
  int rc = gather_frame(ctx);
  if (rc) return rc;
  rc = ensure(ctx, (void**)&ctx->fb.rgb, &ctx->fb.rgb_bytes, npix * 3 + 16);
  if (rc) return rc;
  WX_CUDA(ctx, launch_srgb_rgb8(ctx->fb.rgba, ctx->fb.rgb, npix, d0.stream));
  WX_CUDA(ctx, cudaMemcpyAsync(rgb_out, ctx->fb.rgb, npix * 3, cudaMemcpyDeviceToHost, d0.stream));
  WX_CUDA(ctx, cudaStreamSynchronize(d0.stream));
  return WX_OK;
} catch (const std::bad_alloc&) {
  return fail(nullptr, WX_ERR_OUT_OF_MEMORY, "wx_capture_srgb: host allocation failed");
} catch (const std::exception& ex) {
  return fail(nullptr, WX_ERR_UNSUPPORTED, std::string("wx_capture_srgb: ") + ex.what());
} catch (...) {
  return fail(nullptr, WX_ERR_UNSUPPORTED, "wx_capture_srgb: unexpected exception");
}

extern "C" int wx_last_render_info(const WxContext* cctx, WxRenderInfo* info) try {
  WxContext* ctx = const_cast<WxContext*>(cctx);
  if (!ctx || !info) return WX_ERR_INVALID_ARGUMENT;
  float kmax = 0.f;
  for (DeviceSlot& s : ctx->dev) {
    if (!s.events_pending) continue;
    WX_CUDA(ctx, cudaSetDevice(s.id));
    WX_CUDA(ctx, cudaEventSynchronize(s.ev1));
    float ms = 0.f;
    WX_CUDA(ctx, cudaEventElapsedTime(&ms, s.ev0, s.ev1));
    kmax = std::max(kmax, ms);
  }
  ctx->info.kernel_ms = kmax;
  if (ctx->total_pending) {
    WX_CUDA(ctx, cudaSetDevice(ctx->dev[0].id));
    WX_CUDA(ctx, cudaEventSynchronize(ctx->total1));
    float ms = 0.f;
    WX_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->total0, ctx->total1));
    ctx->info.total_ms = ms;
  }
  (void)cudaSetDevice(ctx->dev[0].id);
  *info = ctx->info;
  return WX_OK;
} catch (const std::bad_alloc&) {
  return fail(nullptr, WX_ERR_OUT_OF_MEMORY, "wx_last_render_info: host allocation failed");
} catch (const std::exception& ex) {
  return fail(nullptr, WX_ERR_UNSUPPORTED, std::string("wx_last_render_info: ") + ex.what());
} catch (...) {
  return fail(nullptr, WX_ERR_UNSUPPORTED, "wx_last_render_info: unexpected exception");
}

// ---------------------------------------------------------------------------------------------
// Memory helpers
// ---------------------------------------------------------------------------------------------
static int slot_of(WxContext* ctx, int device_index, DeviceSlot** s) {
  if (!ctx || device_index < 0 || device_index >= (int)ctx->dev.size()) return fail(ctx, WX_ERR_INVALID_ARGUMENT, "bad device index");
  *s = &ctx->dev[device_index];
  WX_CUDA(ctx, cudaSetDevice((*s)->id));
  return WX_OK;
}

extern "C" int wx_device_alloc(WxContext* ctx, int device_index, size_t bytes, void** out) {
  DeviceSlot* s;
  if (!out) return WX_ERR_INVALID_ARGUMENT;
  int rc = slot_of(ctx, device_index, &s);
  if (rc) return rc;
  WX_CUDA(ctx, cudaMalloc(out, bytes ? bytes : 1));
  return WX_OK;
}
extern "C" int wx_device_free(WxContext* ctx, int device_index, void* ptr) {
  DeviceSlot* s;
  int rc = slot_of(ctx, device_index, &s);
  if (rc) return rc;
  WX_CUDA(ctx, cudaFree(ptr));
  return WX_OK;
}
extern "C" int wx_host_alloc_pinned(size_t bytes, void** out) {
  if (!out) return WX_ERR_INVALID_ARGUMENT;
  cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable);
  if (e != cudaSuccess) return fail_cuda(nullptr, e, "cudaHostAlloc");
  return WX_OK;
}
extern "C" int wx_host_free_pinned(void* ptr) {
  cudaError_t e = cudaFreeHost(ptr);
  if (e != cudaSuccess) return fail_cuda(nullptr, e, "cudaFreeHost");
  return WX_OK;
}
extern "C" int wx_memcpy_d2h(WxContext* ctx, int device_index, void* dst_host, const void* src_dev, size_t bytes, void* stream) {
  DeviceSlot* s;
  int rc = slot_of(ctx, device_index, &s);
  if (rc) return rc;
  WX_CUDA(ctx, cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, reinterpret_cast<cudaStream_t>(stream)));
  return WX_OK;
}
extern "C" int wx_stream_synchronize(WxContext* ctx, int device_index, void* stream) {
  DeviceSlot* s;
  int rc = slot_of(ctx, device_index, &s);
  if (rc) return rc;
  WX_CUDA(ctx, cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(stream)));
  return WX_OK;
}
extern "C" int wx_ipc_export(WxContext* ctx, int device_index, void* ptr, uint8_t handle_out[64]) {
  DeviceSlot* s;
  int rc = slot_of(ctx, device_index, &s);
  if (rc) return rc;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
  cudaIpcMemHandle_t h;
  WX_CUDA(ctx, cudaIpcGetMemHandle(&h, ptr));
  memcpy(handle_out, &h, 64);
  return WX_OK;
}
extern "C" int wx_ipc_open(WxContext* ctx, int device_index, const uint8_t handle[64], void** out) {
  DeviceSlot* s;
  if (!out) return WX_ERR_INVALID_ARGUMENT;
  int rc = slot_of(ctx, device_index, &s);
  if (rc) return rc;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  WX_CUDA(ctx, cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
  return WX_OK;
}
extern "C" int wx_ipc_close(WxContext* ctx, int device_index, void* ptr) {
  DeviceSlot* s;
  int rc = slot_of(ctx, device_index, &s);
  if (rc) return rc;
  WX_CUDA(ctx, cudaIpcCloseMemHandle(ptr));
  return WX_OK;
}
