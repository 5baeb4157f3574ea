// wx_sdf.cu -- VDB345::compute_sdf (src/vdb/vdb345.rs:290-628) on the GPU, value for value.
//
// The reference runs one forward and one backward chamfer sweep over the tree in DFS order (N5s by
// origin, slots ascending, recursively); every tile / inactive voxel takes min(self, neighbour + 1)
// over the 13 already-visited neighbours of its own level, a neighbour that is a child, an active
// voxel, of another level or missing contributes 1, and the values read are whatever the sweep has
// produced so far (the result depends on the traversal order).  What makes it parallel
// without changing a value:
//   (1) the three levels never read each other's distances, so each level is swept on its own;
//   (2) the value a slot ends a pass with is a pure function of the pass-FINAL values of those of its
//       13 neighbours that precede it in the pass and of the pass-INITIAL values of the others.
//       "Precedes" is lexicographic (x, y, z) order inside a node and the DFS index between nodes, so
//       a pass is the unique solution of a recurrence over a DAG:
//           v(c) = min(init(c), 1 + min over the 13 n of (n precedes c ? v(n) : init(n)))
//       (init = MAX-1 in the forward pass, the forward result in the backward pass; seeds are 0);
//   (3) that solution is also the limit of relaxing all nodes at once from any upper bound: every
//       relaxed value stays an upper bound, and a value d is exact after at most d rounds because its
//       derivation chain has d links.  Distances are small (<= 16 on every scene here), so a pass is
//       a handful of rounds in which every node is independent: one CTA per node loads the node plus
//       a one-cell halo of its 26 neighbours into shared memory, sweeps it exactly in 7(DIM-1)+1
//       wavefronts (4x + 2y + z is smaller for all 13 neighbours) and writes it back; rounds repeat
//       until none changes a value.
// Forward results live in F, backward (final) results in B.
#include <algorithm>
#include <cstring>
#include <vector>

#include "wx_device.cuh"
#include "wx_internal.h"

namespace wx {
namespace sdf {

constexpr int kNoNode = -1;

// ---------------------------------------------------------------------------------------------
// topology helpers on the flat (reference order) arrays
// ---------------------------------------------------------------------------------------------
struct Topo {
  uint32_t n5, n4, n3;
  const int32_t* origins;  // n5 x 3
  const uint64_t* kids5;
  const uint32_t* tab5;    // child index where the kid bit is set
  const uint64_t* kids4;
  const uint32_t* tab4;
  const uint64_t* vals3;
};

__device__ __forceinline__ bool bit64(const uint64_t* m, uint32_t i) { return (m[i >> 6] >> (i & 63)) & 1ull; }

// node of `level` (5, 4, 3) that contains global voxel g, or kNoNode (get_voxel's walk, vdb345.rs:69-106)
__device__ int find_node(const Topo& T, int level, long long gx, long long gy, long long gz) {
  if (gx < -2147483648ll || gx > 2147483647ll || gy < -2147483648ll || gy > 2147483647ll || gz < -2147483648ll || gz > 2147483647ll)
    return kNoNode;
  const int x = (int)gx, y = (int)gy, z = (int)gz;
  const int ox = (x >> 12) << 12, oy = (y >> 12) << 12, oz = (z >> 12) << 12;
  int i5 = kNoNode;
  for (uint32_t i = 0; i < T.n5; ++i)
    if (T.origins[3 * i] == ox && T.origins[3 * i + 1] == oy && T.origins[3 * i + 2] == oz) {
      i5 = (int)i;
      break;
    }
  if (i5 < 0 || level == 5) return i5;
  const uint32_t o5 = (((uint32_t)(x & 4095) >> 7) << 10) | (((uint32_t)(y & 4095) >> 7) << 5) | ((uint32_t)(z & 4095) >> 7);
  if (!bit64(T.kids5 + (size_t)i5 * 512, o5)) return kNoNode;
  const int i4 = (int)T.tab5[(size_t)i5 * 32768 + o5];
  if (level == 4) return i4;
  const uint32_t o4 = (((uint32_t)(x & 127) >> 3) << 8) | (((uint32_t)(y & 127) >> 3) << 4) | ((uint32_t)(z & 127) >> 3);
  if (!bit64(T.kids4 + (size_t)i4 * 64, o4)) return kNoNode;
  return (int)T.tab4[(size_t)i4 * 4096 + o4];
}

// origins of the N4 nodes / leaves from their parents' origins
__global__ void child_origins_kernel(const uint64_t* kids, const uint32_t* tab, const int32_t* parent_org, uint32_t n_parents,
                                     int log2d, int child_edge, int32_t* child_org) {
  const uint32_t slots = 1u << (3 * log2d);
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n_parents * slots) return;
  const uint32_t p = (uint32_t)(i / slots), o = (uint32_t)(i % slots);
  if (!bit64(kids + (size_t)p * (slots / 64), o)) return;
  const uint32_t c = tab[i], m = (1u << log2d) - 1u;
  child_org[3 * c + 0] = parent_org[3 * p + 0] + (int32_t)((o >> (2 * log2d)) & m) * child_edge;
  child_org[3 * c + 1] = parent_org[3 * p + 1] + (int32_t)((o >> log2d) & m) * child_edge;
  child_org[3 * c + 2] = parent_org[3 * p + 2] + (int32_t)(o & m) * child_edge;
}

// nb[k][27]: same-level node at origin + d * node_edge, d in {-1,0,1}^3 (index (dx+1)*9 + (dy+1)*3 + dz+1)
__global__ void neighbours_kernel(Topo T, int level, const int32_t* org, uint32_t n, int node_edge, int32_t* nb) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n * 27) return;
  const uint32_t k = (uint32_t)(i / 27), d = (uint32_t)(i % 27);
  const int dx = (int)(d / 9) - 1, dy = (int)((d / 3) % 3) - 1, dz = (int)(d % 3) - 1;
  nb[i] = d == 13 ? (int32_t)k
                  : find_node(T, level, (long long)org[3 * k] + (long long)dx * node_edge, (long long)org[3 * k + 1] + (long long)dy * node_edge,
                              (long long)org[3 * k + 2] + (long long)dz * node_edge);
}

// ---------------------------------------------------------------------------------------------
// one pass over one level
// ---------------------------------------------------------------------------------------------
template <class V>
struct Inf;
template <>
struct Inf<uint32_t> {
  static constexpr uint32_t v = 0xFFFFFFFEu;  // MAX - 1 "so adding 1 doesn't wrap around" (vdb345.rs:300-319)
};
template <>
struct Inf<uint16_t> {
  static constexpr uint16_t v = 0xFFFEu;  // the same construction in 16 bits (leaf distances; checked on exit)
};

template <class V>
struct LevelPass {
  uint32_t n;              // nodes of the level
  const uint64_t* seeds;   // per node SLOTS/64 words: child mask (internal) / value mask (leaf): these contribute 1
  const int32_t* nb;       // [n][27]
  V* F;                    // forward results  [n][SLOTS]
  V* B;                    // backward results [n][SLOTS]
  uint32_t* changed;       // set when a round lowers a value
  const uint8_t* dirty_prev;  // [n]: the node changed in the previous round (nullptr: first round of a pass, sweep everything)
  uint8_t* dirty_cur;         // [n]: set when this round lowers a value of the node
  int backward;
};

template <class V>
__device__ __forceinline__ V ld_cg(const V* p);
template <>
__device__ __forceinline__ uint32_t ld_cg<uint32_t>(const uint32_t* p) { return __ldcg(p); }
template <>
__device__ __forceinline__ uint16_t ld_cg<uint16_t>(const uint16_t* p) { return __ldcg(p); }

// LOG2D = 5, 4, 3; DIM^2 threads (one per (x, y) column), dynamic shared memory (DIM+2)^3 * sizeof(V)
template <int LOG2D, class V>
__global__ void sweep_kernel(const LevelPass<V> L) {
  constexpr int DIM = 1 << LOG2D, H = DIM + 2;
  constexpr uint32_t SLOTS = 1u << (3 * LOG2D);
  constexpr V INF = Inf<V>::v;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  V* halo = reinterpret_cast<V*>(smem_raw);  // [H][H][H], index ((x+1)*H + (y+1))*H + (z+1)
  __shared__ int32_t s_nb[27];
  __shared__ uint32_t s_changed;
  const int tid = threadIdx.x;
  const uint32_t k = blockIdx.x;
  if (tid == 0) s_changed = 0u;
  __syncthreads();
  if (tid < 27) {
    const int32_t j = L.nb[(size_t)k * 27 + tid];
    s_nb[tid] = j;
    // a node's sweep reads its own pass-initial values and its neighbours' current ones: it can only produce something
    // new if a neighbour changed in the previous round
    if (L.dirty_prev && j >= 0 && tid != 13 && L.dirty_prev[j]) s_changed = 1u;
  }
  __syncthreads();
  if (L.dirty_prev && !s_changed) return;
  __syncthreads();
  if (tid == 0) s_changed = 0u;
  __syncthreads();

  // ---- fill the halo cube --------------------------------------------------------------------
  V* cur_buf = L.backward ? L.B : L.F;
  for (int c = tid; c < H * H * H; c += blockDim.x) {
    const int hz = c % H, hy = (c / H) % H, hx = c / (H * H);
    const int dx = hx == 0 ? -1 : (hx == H - 1 ? 1 : 0), dy = hy == 0 ? -1 : (hy == H - 1 ? 1 : 0), dz = hz == 0 ? -1 : (hz == H - 1 ? 1 : 0);
    const int32_t j = s_nb[(dx + 1) * 9 + (dy + 1) * 3 + dz + 1];
    V v;
    if (j < 0) {
      v = 0;  // no node of this level there: contributes 1
    } else {
      const uint32_t o = ((uint32_t)((hx - 1) & (DIM - 1)) << (2 * LOG2D)) | ((uint32_t)((hy - 1) & (DIM - 1)) << LOG2D) | (uint32_t)((hz - 1) & (DIM - 1));
      if (bit64(L.seeds + (size_t)j * (SLOTS / 64), o)) {
        v = 0;  // child / active voxel: contributes 1
      } else if ((uint32_t)j == k) {
        v = L.backward ? L.F[(size_t)k * SLOTS + o] : INF;  // own slots start from the previous pass
      } else {
        const bool precedes = L.backward ? (uint32_t)j > k : (uint32_t)j < k;
        if (precedes) v = ld_cg(cur_buf + (size_t)j * SLOTS + o);           // this pass' value so far (an upper bound; exact at the fixed point)
        else v = L.backward ? L.F[(size_t)j * SLOTS + o] : INF;             // follows this node: its pass-initial value
      }
    }
    halo[c] = v;
  }
  __syncthreads();

  // ---- wavefronts: slot (x, y, z) of the pass-oriented node in front 4x + 2y + z --------------
  const int px = tid >> LOG2D, py = tid & (DIM - 1);  // pass-oriented column of this thread
  const int sgn = L.backward ? -1 : 1;
  for (int w = 0; w <= 7 * (DIM - 1); ++w) {
    const int pz = w - 4 * px - 2 * py;
    if (tid < DIM * DIM && pz >= 0 && pz < DIM) {
      // real coordinates: the backward pass walks the node from its far corner
      const int x = L.backward ? DIM - 1 - px : px, y = L.backward ? DIM - 1 - py : py, z = L.backward ? DIM - 1 - pz : pz;
      const uint32_t o = ((uint32_t)x << (2 * LOG2D)) | ((uint32_t)y << LOG2D) | (uint32_t)z;
      if (!bit64(L.seeds + (size_t)k * (SLOTS / 64), o)) {
        const int c = ((x + 1) * H + (y + 1)) * H + (z + 1);
        V cur = halo[c];
        // the 13 neighbours of vdb345.rs:327-343 (forward: (-1,*,*), (0,-1,*), (0,0,-1); backward: negated)
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
          for (int dz = -1; dz <= 1; ++dz) {
            const V s = halo[c - sgn * H * H + dy * H + dz];
            cur = min(cur, (V)(s + 1));
          }
#pragma unroll
        for (int dz = -1; dz <= 1; ++dz) {
          const V s = halo[c - sgn * H + dz];
          cur = min(cur, (V)(s + 1));
        }
        {
          const V s = halo[c - sgn];
          cur = min(cur, (V)(s + 1));
        }
        halo[c] = cur;
      }
    }
    __syncthreads();
  }

  // ---- write back; report whether this round still lowered something ------------------------------
  bool lowered = false;
  for (uint32_t o = tid; o < SLOTS; o += blockDim.x) {
    const int x = (int)(o >> (2 * LOG2D)), y = (int)((o >> LOG2D) & (DIM - 1)), z = (int)(o & (DIM - 1));
    const V v = halo[((x + 1) * H + (y + 1)) * H + (z + 1)];
    V* dst = cur_buf + (size_t)k * SLOTS + o;
    if (*dst != v) {
      *dst = v;
      lowered = true;
    }
  }
  if (lowered) s_changed = 1u;
  __syncthreads();
  if (tid == 0 && s_changed) {
    *L.changed = 1u;
    L.dirty_cur[k] = 1;
  }
}

template <class V>
__global__ void fill_kernel(V* p, size_t n, V v) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// distances -> the reference's table: child index where the child bit is set, else the distance (u32)
__global__ void merge_internal_kernel(const uint64_t* kids, const uint32_t* tab_in, const uint32_t* dist, size_t total, uint32_t slots_log2,
                                      uint32_t* tab_out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const size_t node = i >> slots_log2;
  const uint32_t o = (uint32_t)(i & ((1u << slots_log2) - 1u));
  tab_out[i] = bit64(kids + node * ((1u << slots_log2) / 64), o) ? tab_in[i] : dist[i];
}

// leaf distances: 0 where the voxel is active; u8 or u32 out; *bad counts values that do not fit
template <class OUT>
__global__ void merge_leaf_kernel(const uint64_t* vals3, const uint16_t* dist, size_t total, OUT* out, uint32_t* max_seen, uint32_t* bad) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const size_t node = i >> 9;
  const uint32_t o = (uint32_t)(i & 511u);
  uint32_t d = 0, m = 0, b = 0;
  if (!bit64(vals3 + node * 8, o)) {
    d = dist[i];
    if (d == 0xFFFEu) d = 0xFFFFFFFEu;  // never reached: the reference leaves MAX - 1
    else if (d >= 0x8000u) b = 1;       // too close to the 16-bit sentinel to be trusted
    if (d != 0xFFFFFFFEu) m = d;
  }
  if (sizeof(OUT) == 1 && d > 255u) b = 1;
  out[i] = (OUT)d;
  // one atomic per warp (the grid is a multiple of 32 threads; tail threads returned above only in the last warp)
  const uint32_t act = __activemask();
  m = __reduce_max_sync(act, m), b = __reduce_add_sync(act, b);
  if ((threadIdx.x & 31u) == (uint32_t)(__ffs(act) - 1)) {
    if (m) atomicMax(max_seen, m);
    if (b) atomicAdd(bad, b);
  }
}

// distances -> the raycast kernel's tables, on the device (wx_tree_build): the packing wx_tree_upload does on the host
// (wx_device.cuh, DevTree): active tile -> 0.0f, child -> flag | index, tile -> f32 bits of dist * cell
__global__ void pack_internal_kernel(const uint64_t* kids, const uint64_t* vals, const uint32_t* tab_in, const uint32_t* dist, size_t total,
                                     uint32_t slots_log2, float cell, uint32_t* e_out, uint32_t* max_seen) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const size_t node = i >> slots_log2;
  const uint32_t o = (uint32_t)(i & ((1u << slots_log2) - 1u)), words = (1u << slots_log2) / 64;
  uint32_t e;
  if (bit64(vals + node * words, o)) {
    e = 0u;
  } else if (bit64(kids + node * words, o)) {
    e = kChildFlag | tab_in[i];
  } else {
    const uint32_t dd = dist[i];
    e = __float_as_uint((float)dd * cell);
    atomicMax(max_seen, dd);
  }
  e_out[i] = e;
}
__global__ void pack_leaf_kernel(const uint64_t* vals3, const uint16_t* dist, size_t total, uint8_t* l3_out, uint32_t* max_seen, uint32_t* bad) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  uint32_t d = 0, m = 0, b = 0;
  if (!bit64(vals3 + (i >> 9) * 8, (uint32_t)(i & 511u))) {
    d = dist[i];
    if (d > 255u) b = 1;  // needs the u32 brick layout: the caller falls back to wx_compute_sdf + wx_tree_upload
    else m = d;
  }
  l3_out[i] = (uint8_t)d;
  const uint32_t act = __activemask();
  m = __reduce_max_sync(act, m), b = __reduce_add_sync(act, b);
  if ((threadIdx.x & 31u) == (uint32_t)(__ffs(act) - 1)) {
    if (m) atomicMax(max_seen, m);
    if (b) atomicAdd(bad, b);
  }
}

}  // namespace sdf

// ---------------------------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------------------------
#define SDF_CUDA(call)                          \
  do {                                          \
    cudaError_t e_ = (call);                    \
    if (e_ != cudaSuccess) {                    \
      cleanup();                                \
      return e_;                                \
    }                                           \
  } while (0)

// One pass of one level: rounds of the relaxation kernel until a round changes nothing.
template <int LOG2D, class V>
static cudaError_t run_pass(sdf::LevelPass<V> L, cudaStream_t stream, uint32_t* rounds) {
  constexpr int DIM = 1 << LOG2D, H = DIM + 2;
  constexpr size_t SLOTS = (size_t)1 << (3 * LOG2D);
  const size_t smem = (size_t)H * H * H * sizeof(V);
  const int threads = std::max(DIM * DIM, 32);
  if (smem > 48 * 1024) {  // per device (and cheap): set on every call, so that a context on another GPU gets the opt-in too
    cudaError_t e = cudaFuncSetAttribute(sdf::sweep_kernel<LOG2D, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  if (L.n == 0) return cudaSuccess;
  const size_t cells = (size_t)L.n * SLOTS;
  cudaError_t e;
  // upper bound the relaxation starts from: MAX-1 (forward), the forward result (backward)
  if (L.backward) {
    e = cudaMemcpyAsync(L.B, L.F, cells * sizeof(V), cudaMemcpyDeviceToDevice, stream);
  } else {
    sdf::fill_kernel<V><<<(unsigned)((cells + 255) / 256), 256, 0, stream>>>(L.F, cells, sdf::Inf<V>::v);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) return e;
  uint8_t* dirty[2] = {L.dirty_cur, const_cast<uint8_t*>(L.dirty_prev)};  // the caller passes two scratch arrays of n bytes
  for (uint32_t r = 0; r < 100000u; ++r) {
    e = cudaMemsetAsync(L.changed, 0, sizeof(uint32_t), stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(dirty[r & 1], 0, L.n, stream);
    if (e != cudaSuccess) return e;
    L.dirty_cur = dirty[r & 1];
    L.dirty_prev = r == 0 ? nullptr : dirty[(r & 1) ^ 1];
    sdf::sweep_kernel<LOG2D, V><<<L.n, threads, smem, stream>>>(L);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    uint32_t changed = 0;
    e = cudaMemcpyAsync(&changed, L.changed, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) return e;
    ++*rounds;
    if (!changed) break;
  }
  return cudaSuccess;
}

// Everything on `stream` of the current device.  Host arrays in, host arrays out (tab3_out: u8 when
// tab3_elem_bytes == 1, else u32).  info: [0..2] max distance per level, [3] values that did not fit, [4] rounds.
cudaError_t compute_sdf_device(const WxTreeDesc& d, uint32_t* tab5_out, uint32_t* tab4_out, void* tab3_out, uint32_t tab3_elem_bytes,
                               uint32_t info[5], float* device_ms, cudaStream_t stream, const SdfDeviceTargets* dev) {
  std::vector<void*> allocs;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  auto cleanup = [&]() {
    for (void* p : allocs) (void)cudaFree(p);
    if (ev0) (void)cudaEventDestroy(ev0);
    if (ev1) (void)cudaEventDestroy(ev1);
  };
  auto dalloc = [&](void** p, size_t bytes) -> cudaError_t {
    cudaError_t e = cudaMalloc(p, bytes ? bytes : 16);
    if (e == cudaSuccess) allocs.push_back(*p);
    return e;
  };
  auto upload = [&](void** p, const void* src, size_t bytes) -> cudaError_t {
    cudaError_t e = dalloc(p, bytes);
    if (e == cudaSuccess && bytes) e = cudaMemcpyAsync(*p, src, bytes, cudaMemcpyHostToDevice, stream);
    return e;
  };
  const size_t s5 = (size_t)d.n5 * 32768, s4 = (size_t)d.n4 * 4096, s3 = (size_t)d.n3 * 512;
  int32_t *org5, *org4, *org3, *nb5, *nb4, *nb3;
  uint64_t *kids5, *kids4, *vals3, *vals5 = nullptr, *vals4 = nullptr;
  uint32_t *tab5, *tab4, *F5, *B5, *F4, *B4, *misc, *out5 = nullptr, *out4 = nullptr;
  uint16_t *F3, *B3;
  void* out3 = nullptr;
  SDF_CUDA(cudaEventCreate(&ev0));
  SDF_CUDA(cudaEventCreate(&ev1));
  SDF_CUDA(upload((void**)&org5, d.origins, (size_t)d.n5 * 12));
  SDF_CUDA(upload((void**)&kids5, d.kids5, (size_t)d.n5 * 4096));
  SDF_CUDA(upload((void**)&tab5, d.tab5, s5 * 4));
  SDF_CUDA(upload((void**)&kids4, d.kids4, (size_t)d.n4 * 512));
  SDF_CUDA(upload((void**)&tab4, d.tab4, s4 * 4));
  SDF_CUDA(upload((void**)&vals3, d.vals3, (size_t)d.n3 * 64));
  if (dev) {
    SDF_CUDA(upload((void**)&vals5, d.vals5, (size_t)d.n5 * 4096));
    SDF_CUDA(upload((void**)&vals4, d.vals4, (size_t)d.n4 * 512));
  }
  SDF_CUDA(dalloc((void**)&org4, (size_t)d.n4 * 12));
  SDF_CUDA(dalloc((void**)&org3, (size_t)d.n3 * 12));
  SDF_CUDA(dalloc((void**)&nb5, (size_t)d.n5 * 27 * 4));
  SDF_CUDA(dalloc((void**)&nb4, (size_t)d.n4 * 27 * 4));
  SDF_CUDA(dalloc((void**)&nb3, (size_t)d.n3 * 27 * 4));
  SDF_CUDA(dalloc((void**)&F5, s5 * 4));
  SDF_CUDA(dalloc((void**)&B5, s5 * 4));
  SDF_CUDA(dalloc((void**)&F4, s4 * 4));
  SDF_CUDA(dalloc((void**)&B4, s4 * 4));
  SDF_CUDA(dalloc((void**)&F3, s3 * 2));
  SDF_CUDA(dalloc((void**)&B3, s3 * 2));
  if (!dev) {
    SDF_CUDA(dalloc((void**)&out5, s5 * 4));
    SDF_CUDA(dalloc((void**)&out4, s4 * 4));
    SDF_CUDA(dalloc(&out3, s3 * (tab3_elem_bytes == 1 ? 1 : 4)));
  }
  uint8_t *dirty_a, *dirty_b;
  const size_t n_max = std::max<size_t>(std::max<size_t>(d.n5, d.n4), d.n3);
  SDF_CUDA(dalloc((void**)&dirty_a, n_max));
  SDF_CUDA(dalloc((void**)&dirty_b, n_max));
  SDF_CUDA(dalloc((void**)&misc, 64));  // [0] changed flag, [4..6] max distance per level, [7] values that do not fit
  SDF_CUDA(cudaMemsetAsync(misc, 0, 64, stream));
  SDF_CUDA(cudaEventRecord(ev0, stream));

  sdf::Topo T{d.n5, d.n4, d.n3, org5, kids5, tab5, kids4, tab4, vals3};
  auto blocks = [](size_t n) { return (unsigned)((n + 255) / 256); };
  if (s5) sdf::child_origins_kernel<<<blocks(s5), 256, 0, stream>>>(kids5, tab5, org5, d.n5, 5, 128, org4);
  if (s4) sdf::child_origins_kernel<<<blocks(s4), 256, 0, stream>>>(kids4, tab4, org4, d.n4, 4, 8, org3);
  if (d.n5) sdf::neighbours_kernel<<<blocks((size_t)d.n5 * 27), 256, 0, stream>>>(T, 5, org5, d.n5, 4096, nb5);
  if (d.n4) sdf::neighbours_kernel<<<blocks((size_t)d.n4 * 27), 256, 0, stream>>>(T, 4, org4, d.n4, 128, nb4);
  if (d.n3) sdf::neighbours_kernel<<<blocks((size_t)d.n3 * 27), 256, 0, stream>>>(T, 3, org3, d.n3, 8, nb3);
  SDF_CUDA(cudaGetLastError());

  uint32_t rounds = 0;
  for (int pass = 0; pass < 2; ++pass) {
    sdf::LevelPass<uint32_t> L5{d.n5, kids5, nb5, F5, B5, misc, dirty_b, dirty_a, pass};
    SDF_CUDA((run_pass<5, uint32_t>(L5, stream, &rounds)));
    sdf::LevelPass<uint32_t> L4{d.n4, kids4, nb4, F4, B4, misc, dirty_b, dirty_a, pass};
    SDF_CUDA((run_pass<4, uint32_t>(L4, stream, &rounds)));
    sdf::LevelPass<uint16_t> L3{d.n3, vals3, nb3, F3, B3, misc, dirty_b, dirty_a, pass};
    SDF_CUDA((run_pass<3, uint16_t>(L3, stream, &rounds)));
  }
  uint32_t h_misc[16];
  uint32_t m5 = 0, m4 = 0;
  if (dev) {  // straight into the raycast kernel's tables
    if (s5) sdf::pack_internal_kernel<<<blocks(s5), 256, 0, stream>>>(kids5, vals5, tab5, B5, s5, 15, 128.f, dev->e5, misc + 4);
    if (s4) sdf::pack_internal_kernel<<<blocks(s4), 256, 0, stream>>>(kids4, vals4, tab4, B4, s4, 12, 8.f, dev->e4, misc + 5);
    if (s3) sdf::pack_leaf_kernel<<<blocks(s3), 256, 0, stream>>>(vals3, B3, s3, dev->l3, misc + 6, misc + 7);
    SDF_CUDA(cudaGetLastError());
    SDF_CUDA(cudaEventRecord(ev1, stream));
    SDF_CUDA(cudaMemcpyAsync(h_misc, misc, 64, cudaMemcpyDeviceToHost, stream));
    SDF_CUDA(cudaStreamSynchronize(stream));
    if (device_ms) SDF_CUDA(cudaEventElapsedTime(device_ms, ev0, ev1));
    m5 = h_misc[4], m4 = h_misc[5];
  } else {
    if (s5) sdf::merge_internal_kernel<<<blocks(s5), 256, 0, stream>>>(kids5, tab5, B5, s5, 15, out5);
    if (s4) sdf::merge_internal_kernel<<<blocks(s4), 256, 0, stream>>>(kids4, tab4, B4, s4, 12, out4);
    if (s3) {
      if (tab3_elem_bytes == 1) sdf::merge_leaf_kernel<uint8_t><<<blocks(s3), 256, 0, stream>>>(vals3, B3, s3, (uint8_t*)out3, misc + 6, misc + 7);
      else sdf::merge_leaf_kernel<uint32_t><<<blocks(s3), 256, 0, stream>>>(vals3, B3, s3, (uint32_t*)out3, misc + 6, misc + 7);
    }
    SDF_CUDA(cudaGetLastError());
    SDF_CUDA(cudaEventRecord(ev1, stream));
    if (s5) SDF_CUDA(cudaMemcpyAsync(tab5_out, out5, s5 * 4, cudaMemcpyDeviceToHost, stream));
    if (s4) SDF_CUDA(cudaMemcpyAsync(tab4_out, out4, s4 * 4, cudaMemcpyDeviceToHost, stream));
    if (s3) SDF_CUDA(cudaMemcpyAsync(tab3_out, out3, s3 * (tab3_elem_bytes == 1 ? 1 : 4), cudaMemcpyDeviceToHost, stream));
    SDF_CUDA(cudaMemcpyAsync(h_misc, misc, 64, cudaMemcpyDeviceToHost, stream));
    SDF_CUDA(cudaStreamSynchronize(stream));
    if (device_ms) SDF_CUDA(cudaEventElapsedTime(device_ms, ev0, ev1));
    // max tile distances of the internal levels: on the host, from the merged tables
    for (size_t i = 0; i < s5; ++i)
      if (!((d.kids5[i >> 6] >> (i & 63)) & 1ull) && tab5_out[i] != 0xFFFFFFFEu) m5 = std::max(m5, tab5_out[i]);
    for (size_t i = 0; i < s4; ++i)
      if (!((d.kids4[i >> 6] >> (i & 63)) & 1ull) && tab4_out[i] != 0xFFFFFFFEu) m4 = std::max(m4, tab4_out[i]);
  }
  info[0] = m5, info[1] = m4, info[2] = h_misc[6], info[3] = h_misc[7], info[4] = rounds;
  cleanup();
  return cudaSuccess;
}

}  // namespace wx
