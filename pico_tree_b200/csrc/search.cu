// search.cu — batched knn / radius / box kernels and their host drivers.
//
// Batch entry points replace the per-query loops of the reference
// (src/pyco_tree/pico_tree/_pyco_tree/kd_tree.hpp:117-268, examples/benchmark/bm_pico_kd_tree.cpp:63-78).
// Two traversal families (traverse.cuh / traverse_warp.cuh):
//   * thread-per-query for sdim <= 3 and k <= 32: queries are Z-ordered first, so the 32
//     threads of a warp walk almost the same root-to-leaf path — node and leaf loads become
//     warp-wide broadcasts served by L1/L2;
//   * warp-per-query for everything else (any sdim, any k), also selectable with
//     PICO_B200_WARP_PER_QUERY for comparison.
#include <cub/cub.cuh>
#include <sys/mman.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "traverse_warp.cuh"

namespace pico {
namespace {

constexpr int kThreadsPerBlock = 128;
constexpr int kWarpsPerBlock = 8;
constexpr size_t kThreadKMax = 32;  // largest k of the thread-per-query kernels (register k-list)

// ------------------------------------------------------------------ query ordering
// 30-bit (3 x 10) Morton code of the query inside the tree's root box; queries outside are
// clamped. Only the first min(sdim, 3) coordinates take part.
__device__ __forceinline__ uint32_t spread10(uint32_t v) {
  v &= 0x3ffu;
  v = (v | (v << 16)) & 0x030000ffu;
  v = (v | (v << 8)) & 0x0300f00fu;
  v = (v | (v << 4)) & 0x030c30c3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}

// The cell of a coordinate is computed in the tree's scalar type with one multiply-add per coordinate
// (the order only has to be the same for every query of a batch; it never decides a result).
template <typename T>
__global__ void morton_kernel(const T* __restrict__ q, size_t stride, uint32_t nq, int dims, T lo0, T lo1, T lo2,
                              T inv0, T inv1, T inv2, uint32_t* __restrict__ codes, uint32_t* __restrict__ ids) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  const T* p = q + (size_t)i * stride;
  const T l[3] = {lo0, lo1, lo2}, s[3] = {inv0, inv1, inv2};
  uint32_t code = 0;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    if (j < dims) {
      T f = (p[j] - l[j]) * s[j];
      f = f < T(0) ? T(0) : (f > T(1023) ? T(1023) : f);  // (a NaN coordinate converts to cell 0)
      code |= spread10((uint32_t)f) << j;
    }
  }
  codes[i] = code;
  ids[i] = i;
}

// ------------------------------------------------------------------ thread-per-query kernels
template <typename T>
struct KnnArgs {
  const typename NodeOf<T>::type* nodes;
  const T* outer;  // {left_min, right_max} per node, topological metrics only
  const uint2* spans;  // {first point, count} below each node (row storage), or nullptr
  const typename Vec4Of<T>::type* pts4;
  const T* rows;
  const int32_t* indices;
  const T* q;
  size_t q_stride;
  uint32_t nq;
  const uint32_t* perm;  // nullptr = identity
  Neighbor<T>* out;
  int k, sdim, metric, approx;
  uint32_t n_points;
  T e_inv;
  // deep-tree workspace (GlobalStack) or warp stacks
  void* ws;
  size_t ws_stride, ws_depth;
  // warp kernels: shared memory per warp in scalars (query + offsets [+ leaf tile]) and the rows
  // of the staged leaf tile (0 = points are read straight from global memory)
  int warp_smem, tile_rows, cache_rows;
  // warp kernels are persistent (one wave of resident blocks): every warp draws its next query here
  unsigned long long* counter;
  // exact nn over the search image (nn_fat_kernel): collapsed nodes, the array far children are walked in,
  // and the list of queries whose best distance is attained twice (re-run by knn_thread_kernel, which
  // then takes its query count from *nq_from)
  const typename NodeOf<T>::type* fat;
  const typename NodeOf<T>::type* far_nodes;
  uint32_t* tie_count;
  uint32_t* tie_list;
  const uint32_t* nq_from;
};

template <typename T, int DIM, int KMAX, bool FAST, bool DEEP>
__global__ void __launch_bounds__(kThreadsPerBlock) knn_thread_kernel(KnnArgs<T> a) {
  const size_t total = (size_t)gridDim.x * blockDim.x;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t nq = a.nq_from ? *a.nq_from : a.nq;
  for (size_t slot = tid; slot < nq; slot += total) {
    const uint32_t qi = a.perm ? a.perm[slot] : (uint32_t)slot;
    T q[DIM];
    const T* qp = a.q + (size_t)qi * a.q_stride;
#pragma unroll
    for (int j = 0; j < DIM; ++j) q[j] = qp[j];
    Neighbor<T>* out = a.out + (size_t)qi * a.k;
    if (KMAX == 1) {
      VisitNn<T> vis;
      if (DEEP) {
        GlobalStack<T, DIM> st;
        st.stride = a.ws_stride;
        st.depth = a.ws_depth;
        st.node = static_cast<uint32_t*>(a.ws) + tid;
        st.dist = reinterpret_cast<T*>(static_cast<uint32_t*>(a.ws) + a.ws_stride * a.ws_depth) + tid;
        st.off = st.dist + a.ws_stride * a.ws_depth;
        traverse_packed<T, DIM, FAST, kPrimeFirstLeaf>(a.nodes, a.pts4, a.outer, q, a.metric, a.approx != 0, a.e_inv, st,
                                                       vis);
      } else {
        LocalStack<T, DIM, kLocalStack> st;
        traverse_packed<T, DIM, FAST, kPrimeFirstLeaf>(a.nodes, a.pts4, a.outer, q, a.metric, a.approx != 0, a.e_inv, st,
                                                       vis);
      }
      store_neighbor(out, vis.idx, vis.best);
    } else {
      VisitKnn<T, KMAX> vis;
      vis.init(a.k);
      if (DEEP) {
        GlobalStack<T, DIM> st;
        st.stride = a.ws_stride;
        st.depth = a.ws_depth;
        st.node = static_cast<uint32_t*>(a.ws) + tid;
        st.dist = reinterpret_cast<T*>(static_cast<uint32_t*>(a.ws) + a.ws_stride * a.ws_depth) + tid;
        st.off = st.dist + a.ws_stride * a.ws_depth;
        traverse_packed<T, DIM, FAST, kPrimeNone>(a.nodes, a.pts4, a.outer, q, a.metric, a.approx != 0, a.e_inv, st, vis);
      } else {
        LocalStack<T, DIM, kLocalStack> st;
        traverse_packed<T, DIM, FAST, kPrimeBound>(a.nodes, a.pts4, a.outer, q, a.metric, a.approx != 0, a.e_inv, st, vis,
                                                   (int)a.n_points, a.k);
      }
      vis.store(out);
    }
  }
}

// Exact nn (traverse_nn): metric_l2_squared, k = 1, trees no deeper than the local stack. FAT: the first
// descent and the second walk run over the search image (fat.cu); queries with a tie at the best distance are
// listed for the order-exact kernel above.
template <typename T, int DIM, int NREC, bool FAT, int MINB>
__global__ void __launch_bounds__(kThreadsPerBlock, MINB) nn_kernel(KnnArgs<T> a) {
  __shared__ uint32_t s_tag[kSharedSlots][kThreadsPerBlock];
  __shared__ T s_x[kSharedSlots][kThreadsPerBlock];
  __shared__ T s_y[kSharedSlots][kThreadsPerBlock];
  const size_t total = (size_t)gridDim.x * blockDim.x;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (size_t slot = tid; slot < a.nq; slot += total) {
    const uint32_t qi = a.perm ? a.perm[slot] : (uint32_t)slot;
    T q[DIM];
    const T* qp = a.q + (size_t)qi * a.q_stride;
#pragma unroll
    for (int j = 0; j < DIM; ++j) q[j] = qp[j];
    VisitNnTie<T> vis;
    SlotStack<T, kThreadsPerBlock> st;
    st.tag = s_tag;
    st.x = s_x;
    st.y = s_y;
    traverse_nn<T, DIM, NREC, FAT>(a.fat, a.far_nodes, a.pts4, q, st, vis);
    Neighbor<T>* out = a.out + qi;
    store_neighbor(out, vis.idx, vis.best);
    if (FAT && vis.tie) a.tie_list[atomicAdd(a.tie_count, 1u)] = qi;
  }
}

// ------------------------------------------------------------------ isolated leaf scan (measurement)
// SURVEY.md §8d asks for the leaf scan (kd_tree_search.hpp:54-59) timed on its own next to the fused
// traversal: first_leaf_kernel walks root -> first leaf (kd_tree_search.hpp:60-88, no far children) and
// stores the leaf's point range per slot; leaf_scan_kernel then only streams those contiguous float4
// records through the search_nn visitor. Same query order (Z-order slots), same loads, same arithmetic
// as the fused kernel's leaf loop.
template <typename T, int DIM>
__global__ void __launch_bounds__(kThreadsPerBlock) first_leaf_kernel(KnnArgs<T> a, int2* __restrict__ ranges,
                                                                      unsigned long long* __restrict__ n_streamed) {
  const size_t total = (size_t)gridDim.x * blockDim.x;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long mine = 0;
  for (size_t slot = tid; slot < a.nq; slot += total) {
    const uint32_t qi = a.perm ? a.perm[slot] : (uint32_t)slot;
    T q[DIM];
    const T* qp = a.q + (size_t)qi * a.q_stride;
#pragma unroll
    for (int j = 0; j < DIM; ++j) q[j] = qp[j];
    uint32_t node = 0, right, sd;
    T na, nb;
    int lb, le;
    load_node(a.nodes, node, na, nb, right, sd, lb, le);
    while (sd != PICO_B200_LEAF) {
      T v = q[0];
#pragma unroll
      for (int j = 1; j < DIM; ++j)
        if (sd == (uint32_t)j) v = q[j];
      bool go_left;
      T unused;
      branch_choice((int)PICO_B200_METRIC_L2_SQUARED, a.outer, node, na, nb, v, sd, go_left, unused);
      node = go_left ? node + 1 : right;
      load_node(a.nodes, node, na, nb, right, sd, lb, le);
    }
    ranges[slot] = make_int2(lb, le);
    mine += (unsigned long long)(le - lb);
  }
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, o);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(n_streamed, mine);
}

template <typename T, int DIM>
__global__ void __launch_bounds__(kThreadsPerBlock) leaf_scan_kernel(KnnArgs<T> a, const int2* __restrict__ ranges) {
  const size_t total = (size_t)gridDim.x * blockDim.x;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (size_t slot = tid; slot < a.nq; slot += total) {
    const uint32_t qi = a.perm ? a.perm[slot] : (uint32_t)slot;
    T q[DIM];
    const T* qp = a.q + (size_t)qi * a.q_stride;
#pragma unroll
    for (int j = 0; j < DIM; ++j) q[j] = qp[j];
    const int2 r = ranges[slot];
    VisitNn<T> vis;
    for (int i = r.x; i < r.y; ++i) {
      const typename Vec4Of<T>::type p = ldg4(a.pts4 + i);
      T d = metric_first((int)PICO_B200_METRIC_L2_SQUARED, q[0], p.x);
      if (DIM > 1) d = metric_fold((int)PICO_B200_METRIC_L2_SQUARED, d, q[DIM > 1 ? 1 : 0], p.y, 1);
      if (DIM > 2) d = metric_fold((int)PICO_B200_METRIC_L2_SQUARED, d, q[DIM > 2 ? 2 : 0], p.z, 2);
      vis.visit(index_of(p), d);
    }
    store_neighbor(a.out + qi, vis.idx, vis.best);
  }
}

template <typename T>
struct RadiusArgs {
  KnnArgs<T> base;        // out unused
  T radius;               // already scaled by 1/e for the approximate visitor
  uint32_t* counts;       // pass 1: per-query hit count
  const uint64_t* offsets;  // pass 2: exclusive scan of counts
  Neighbor<T>* hits;      // pass 2: packed results
};

template <typename T, int DIM, bool FILL, bool DEEP>
__global__ void __launch_bounds__(kThreadsPerBlock, 16) radius_thread_kernel(RadiusArgs<T> r) {
  const KnnArgs<T>& a = r.base;
  const size_t total = (size_t)gridDim.x * blockDim.x;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (size_t slot = tid; slot < a.nq; slot += total) {
    const uint32_t qi = a.perm ? a.perm[slot] : (uint32_t)slot;
    T q[DIM];
    const T* qp = a.q + (size_t)qi * a.q_stride;
#pragma unroll
    for (int j = 0; j < DIM; ++j) q[j] = qp[j];
    GlobalStack<T, DIM> gst;
    if (DEEP) {
      gst.stride = a.ws_stride;
      gst.depth = a.ws_depth;
      gst.node = static_cast<uint32_t*>(a.ws) + tid;
      gst.dist = reinterpret_cast<T*>(static_cast<uint32_t*>(a.ws) + a.ws_stride * a.ws_depth) + tid;
      gst.off = gst.dist + a.ws_stride * a.ws_depth;
    }
    LocalStack<T, DIM, DEEP ? 1 : kLocalStack> lst;
    if (FILL) {
      VisitRadiusFill<T> vis;
      vis.radius = r.radius;
      vis.begin(r.hits + r.offsets[qi]);
      if (DEEP)
        traverse_packed<T, DIM, false, kPrimeNone>(a.nodes, a.pts4, a.outer, q, a.metric, a.approx != 0, a.e_inv, gst, vis);
      else
        traverse_packed<T, DIM, false, kPrimeNone>(a.nodes, a.pts4, a.outer, q, a.metric, a.approx != 0, a.e_inv, lst, vis);
      vis.finish();
    } else {
      VisitRadiusCount<T> vis;
      vis.radius = r.radius;
      if (DEEP)
        traverse_packed<T, DIM, false, kPrimeNone>(a.nodes, a.pts4, a.outer, q, a.metric, a.approx != 0, a.e_inv, gst, vis);
      else
        traverse_packed<T, DIM, false, kPrimeNone>(a.nodes, a.pts4, a.outer, q, a.metric, a.approx != 0, a.e_inv, lst, vis);
      r.counts[qi] = vis.count;
    }
  }
}

// ------------------------------------------------------------------ warp-per-query kernels
// dynamic shared memory per warp: query[sdim] + offsets[sdim]
template <typename T, bool PACKED, bool REGLIST>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) knn_warp_kernel(KnnArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* smem = reinterpret_cast<T*>(smem_raw);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  T* sq = smem + (size_t)w * a.warp_smem;
  T* so = sq + a.sdim;
  T* tile = so + a.sdim;
  WarpFrame<T>* win = reinterpret_cast<WarpFrame<T>*>(sq + a.warp_smem) - kStackWindow;  // tail of the warp's slice
  const size_t warp_global = (size_t)blockIdx.x * kWarpsPerBlock + w;
  WarpFrame<T>* stack = static_cast<WarpFrame<T>*>(a.ws) + warp_global * a.ws_depth;
  PointSet<T, PACKED> ps{a.pts4, a.rows, a.indices, a.sdim};
  for (;;) {
    // queries differ a lot in cost (pruning): draw them one at a time instead of striding
    unsigned long long slot = 0;
    if (lane == 0) slot = atomicAdd(a.counter, 1ull);
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if (slot >= a.nq) break;
    const uint32_t qi = a.perm ? a.perm[slot] : (uint32_t)slot;
    const T* qp = a.q + (size_t)qi * a.q_stride;
    __syncwarp();
    for (int j = lane; j < a.sdim; j += 32) {
      sq[j] = qp[j];
      so[j] = T(0);
    }
    __syncwarp();
    Neighbor<T>* row = a.out + (size_t)qi * a.k;
    if (REGLIST) {
      WarpVisitKnn<T, WarpKnnReg<T>> vis;
      vis.list.init(a.k);
      traverse_warp<T, PACKED>(a.nodes, a.outer, ps, sq, so, stack, win, a.metric, a.approx != 0, a.e_inv, vis,
                               tile, a.tile_rows, a.spans, a.cache_rows);
      vis.list.store(row);
    } else {
      WarpVisitKnn<T, WarpKnnMem<T>> vis;
      vis.list.init(row, a.k);
      traverse_warp<T, PACKED>(a.nodes, a.outer, ps, sq, so, stack, win, a.metric, a.approx != 0, a.e_inv, vis,
                               tile, a.tile_rows, a.spans, a.cache_rows);
    }
  }
}

template <typename T, bool PACKED>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) radius_warp_kernel(RadiusArgs<T> r) {
  const KnnArgs<T>& a = r.base;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* smem = reinterpret_cast<T*>(smem_raw);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  T* sq = smem + (size_t)w * a.warp_smem;
  T* so = sq + a.sdim;
  T* tile = so + a.sdim;
  WarpFrame<T>* win = reinterpret_cast<WarpFrame<T>*>(sq + a.warp_smem) - kStackWindow;  // tail of the warp's slice
  const size_t warp_global = (size_t)blockIdx.x * kWarpsPerBlock + w;
  WarpFrame<T>* stack = static_cast<WarpFrame<T>*>(a.ws) + warp_global * a.ws_depth;
  PointSet<T, PACKED> ps{a.pts4, a.rows, a.indices, a.sdim};
  for (;;) {
    // queries differ a lot in cost (pruning): draw them one at a time instead of striding
    unsigned long long slot = 0;
    if (lane == 0) slot = atomicAdd(a.counter, 1ull);
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if (slot >= a.nq) break;
    const uint32_t qi = a.perm ? a.perm[slot] : (uint32_t)slot;
    const T* qp = a.q + (size_t)qi * a.q_stride;
    __syncwarp();
    for (int j = lane; j < a.sdim; j += 32) {
      sq[j] = qp[j];
      so[j] = T(0);
    }
    __syncwarp();
    WarpVisitRadius<T> vis;
    vis.radius = r.radius;
    vis.out = r.hits ? r.hits + r.offsets[qi] : nullptr;
    traverse_warp<T, PACKED>(a.nodes, a.outer, ps, sq, so, stack, win, a.metric, a.approx != 0, a.e_inv, vis, tile,
                             a.tile_rows, a.spans, a.cache_rows);
    if (!r.hits && lane == 0) r.counts[qi] = vis.count;
  }
}

template <typename T>
struct BoxArgs {
  const typename NodeOf<T>::type* nodes;
  const T* outer;
  int metric;
  const typename Vec4Of<T>::type* pts4;
  const T* rows;
  const int32_t* indices;
  const T* root_box;
  const T* mins;
  const T* maxs;
  size_t stride;
  uint32_t nb;
  int sdim;
  uint32_t* counts;
  const uint64_t* offsets;
  int32_t* hits;
  void* ws;  // per warp: BoxFrame[depth] + T saved[depth]
  size_t ws_depth;
};

// dynamic shared memory per warp: qmin[sdim] qmax[sdim] box[2*sdim]
template <typename T, bool PACKED>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) box_warp_kernel(BoxArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* smem = reinterpret_cast<T*>(smem_raw);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  T* qmin = smem + (size_t)w * 4 * a.sdim;
  T* qmax = qmin + a.sdim;
  T* sbox = qmax + a.sdim;
  const size_t warp_global = (size_t)blockIdx.x * kWarpsPerBlock + w;
  const size_t total_warps = (size_t)gridDim.x * kWarpsPerBlock;
  const size_t per_warp = a.ws_depth * (sizeof(BoxFrame) + sizeof(T));
  unsigned char* wsb = static_cast<unsigned char*>(a.ws) + warp_global * per_warp;
  BoxFrame* stack = reinterpret_cast<BoxFrame*>(wsb);
  T* saved = reinterpret_cast<T*>(wsb + a.ws_depth * sizeof(BoxFrame));
  PointSet<T, PACKED> ps{a.pts4, a.rows, a.indices, a.sdim};
  for (size_t bi = warp_global; bi < a.nb; bi += total_warps) {
    __syncwarp();
    for (int j = lane; j < a.sdim; j += 32) {
      qmin[j] = a.mins[bi * a.stride + j];
      qmax[j] = a.maxs[bi * a.stride + j];
    }
    for (int j = lane; j < 2 * a.sdim; j += 32) sbox[j] = a.root_box[j];
    __syncwarp();
    int32_t* out = a.hits ? a.hits + a.offsets[bi] : nullptr;
    const uint32_t c = traverse_box_warp<T, PACKED>(a.nodes, a.outer, a.metric, ps, a.indices, qmin, qmax, sbox, stack, saved, out);
    if (!a.hits && lane == 0) a.counts[bi] = c;
  }
}

// Thread-per-box (traverse_box_thread): euclidean spaces, sdim <= 3, trees no deeper than the local stack. Boxes are
// taken in the Z-order of their lower corners so that the threads of a warp walk neighbouring cells.
template <typename T, int DIM>
__global__ void __launch_bounds__(kThreadsPerBlock) box_thread_kernel(BoxArgs<T> a, const uint32_t* __restrict__ perm) {
  const size_t total = (size_t)gridDim.x * blockDim.x;
  for (size_t slot = (size_t)blockIdx.x * blockDim.x + threadIdx.x; slot < a.nb; slot += total) {
    const uint32_t bi = perm ? perm[slot] : (uint32_t)slot;
    T qmin[DIM], qmax[DIM];
#pragma unroll
    for (int j = 0; j < DIM; ++j) {
      qmin[j] = a.mins[(size_t)bi * a.stride + j];
      qmax[j] = a.maxs[(size_t)bi * a.stride + j];
    }
    int32_t* out = a.hits ? a.hits + a.offsets[bi] : nullptr;
    const uint32_t c = traverse_box_thread<T, DIM>(a.nodes, a.pts4, a.indices, qmin, qmax, a.root_box, out);
    if (!a.hits) a.counts[bi] = c;
  }
}

// ------------------------------------------------------------------ host helpers
// per-thread call configuration (pico_b200_set_stream / pico_b200_profile_*)
struct ThreadCfg {
  cudaStream_t user_stream = nullptr;
  bool has_user_stream = false;
  bool profiling = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> spans;
};
thread_local ThreadCfg g_cfg;

// Streams and events are recycled per host thread: creating a stream plus five events costs tens of
// microseconds, which is visible next to a 0.1 ms batch.
struct StreamSet {
  int device = -1;
  cudaStream_t st = nullptr;
  cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
};
struct StreamCache {
  std::vector<StreamSet> free_sets;
  ~StreamCache() {
    // the CUDA context may already be gone at thread exit; leak instead of calling into it
  }
};
thread_local StreamCache g_stream_cache;

// One high-priority stream per host thread and device: the host pipeline orders its chunks there. The ordering
// step of a chunk is a chain of small kernels (Morton codes, radix-sort passes); at normal priority each of
// them queues behind the thousands of pending blocks of the previous chunks' traversal kernels and the chain
// takes 0.4 ms instead of 0.05 ms (profiles/r1/host_pipeline_sweep.txt: order 1.83 -> 2.27 ms for the last
// chunk). With priority its blocks are placed as soon as any block slot frees up.
struct PriorityStream {
  int device = -1;
  cudaStream_t st = nullptr;
};
thread_local PriorityStream g_hp_stream;
int priority_stream(int device, cudaStream_t* st) {
  if (g_hp_stream.st && g_hp_stream.device == device) {
    *st = g_hp_stream.st;
    return 0;
  }
  int least = 0, greatest = 0;
  PICO_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
  cudaStream_t s = nullptr;
  PICO_CUDA(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, greatest));
  g_hp_stream.device = device;  // (a stream of another device is leaked: a host thread rarely switches devices)
  g_hp_stream.st = s;
  *st = s;
  return 0;
}

struct CallCtx {
  cudaStream_t st = nullptr;
  bool owns_stream = true;
  bool async = false;
  cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  int dev = -1;
  std::vector<void*> async_allocs;
  int init(int device, bool want_async = false) {
    PICO_CUDA(cudaSetDevice(device));
    if (g_cfg.has_user_stream) {
      st = g_cfg.user_stream;
      owns_stream = false;
      async = want_async;
      if (!async)
        for (auto& e : ev) PICO_CUDA(cudaEventCreate(&e));
      return 0;
    }
    dev = device;
    auto& fs = g_stream_cache.free_sets;
    for (size_t i = 0; i < fs.size(); ++i) {
      if (fs[i].device == device) {
        st = fs[i].st;
        for (int j = 0; j < 5; ++j) ev[j] = fs[i].ev[j];
        fs.erase(fs.begin() + (long)i);
        return 0;
      }
    }
    PICO_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    for (auto& e : ev) PICO_CUDA(cudaEventCreate(&e));
    return 0;
  }
  bool timed = true;  // record the five stage events (search stats)
  // host pipeline: resident blocks per SM the thread-per-query traversal of a chunk may take (0 = no cap). Leaving
  // a few block slots free lets the small ordering kernels of the next chunk start at once instead of waiting
  // for traversal blocks to drain (profiles/r2/host_pipeline_timeline.txt)
  int blocks_per_sm_cap = 0;
  bool pipelined = false;  // a chunk of the host pipeline: always the device-wide sort (see knn_batch)
  bool fused_order = false;  // exact nn may order and traverse in one kernel (nn_tile_kernel)
  bool probe_order = false;  // measure how coherent this batch arrives (tile_order_kernel on the side)
  int failed = 0;            // error code of a helper that reports through the context
  int mark(int i) {
    if (!async && timed) PICO_CUDA(cudaEventRecord(ev[i], st));
    return 0;
  }
  // brackets a traversal kernel when the thread is profiling
  int span_begin() {
    if (!g_cfg.profiling) return 0;
    cudaEvent_t a, b;
    PICO_CUDA(cudaEventCreate(&a));
    PICO_CUDA(cudaEventCreate(&b));
    PICO_CUDA(cudaEventRecord(a, st));
    g_cfg.spans.emplace_back(a, b);
    return 0;
  }
  int span_end() {
    if (!g_cfg.profiling) return 0;
    PICO_CUDA(cudaEventRecord(g_cfg.spans.back().second, st));
    return 0;
  }
  int alloc(void** p, size_t bytes) {
    PICO_CUDA(cudaMallocAsync(p, bytes ? bytes : 16, st));
    async_allocs.push_back(*p);
    return 0;
  }
  // stream-ordered free of everything allocated so far (safe right after enqueueing)
  void release() {
    for (void* p : async_allocs) cudaFreeAsync(p, st);
    async_allocs.clear();
  }
  ~CallCtx() {
    for (void* p : async_allocs) cudaFreeAsync(p, st);
    if (!async && (st || !owns_stream)) cudaStreamSynchronize(st);
    if (owns_stream && st) {
      StreamSet s;
      s.device = dev;
      s.st = st;
      for (int j = 0; j < 5; ++j) s.ev[j] = ev[j];
      g_stream_cache.free_sets.push_back(s);
    } else {
      for (auto& e : ev)
        if (e) cudaEventDestroy(e);
    }
  }
};


// ------------------------------------------------------------------ big results: device -> pageable host
// Ragged radius / box results can be gigabytes (cfg3: 760 M neighbours = 6 GB) and land in fresh
// malloc'd memory. A plain cudaMemcpy into pageable memory runs at ~2 GB/s. Instead the device
// streams chunks into two pinned staging buffers while host threads move the previous chunk to its
// destination (that copy is page-fault bound, hence several threads).
constexpr size_t kStageBytes = (size_t)64 << 20;

struct Staging {
  std::mutex mu;
  void* buf[2] = {nullptr, nullptr};
};
Staging g_staging;

// host threads of the big copies (PICO_B200_COPY_THREADS overrides; tuning hook)
unsigned copy_threads() {
  static const unsigned v = [] {
    const char* e = getenv("PICO_B200_COPY_THREADS");
    const int x = e ? atoi(e) : 0;
    if (x >= 1 && x <= 64) return (unsigned)x;
    return std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
  }();
  return v;
}

// ... and of the copy of a multi-gigabyte ragged result out of the staging buffers, where nothing else keeps the
// host busy: into fresh pages 16 threads move 6.1 GB in 255 ms against 300 ms with 8; into the cached block both
// take 200 ms (profiles/r2/radius_e2e_v1.txt)
unsigned result_copy_threads() {
  static const unsigned v = getenv("PICO_B200_COPY_THREADS")
                                ? copy_threads()
                                : std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
  return v;
}

void parallel_memcpy(char* dst, const char* src, size_t bytes, unsigned nt = copy_threads()) {
  if (bytes < ((size_t)4 << 20)) nt = 1;
  const size_t per = (bytes / nt + 4095) & ~(size_t)4095;
  std::vector<std::thread> th;
  for (unsigned i = 1; i < nt; ++i) {
    const size_t off = std::min(bytes, i * per), len = std::min(bytes, (i + 1) * per) - off;
    if (len) th.emplace_back([=] { memcpy(dst + off, src + off, len); });
  }
  memcpy(dst, src, std::min(bytes, per));
  for (auto& t : th) t.join();
}

// Ragged results can be gigabytes of fresh host memory (cfg3 radius: 6.1 GB). Large buffers are aligned to 2 MiB and
// offered to the kernel as transparent huge pages: the copy out of the pinned staging buffers then takes one page
// fault per 2 MiB instead of one per 4 KiB. free() releases them like any malloc'd block (pico_b200_free).
//
// One such block is kept when the caller frees it (release_result, behind pico_b200_free) and handed to the next
// big result that fits: its pages are already mapped, so the copy is not slowed by the kernel zeroing 2 MiB at
// every first touch — the reference's users run the same radius search frame after frame. At most
// PICO_B200_RESULT_CACHE_MB (default 8192, 0 = off) stay cached.
struct ResultCache {
  std::mutex mu;
  std::vector<std::pair<void*, size_t>> big;  // every live huge-page block this library handed out (or caches)
  void* spare = nullptr;
  size_t spare_cap = 0;
};
ResultCache g_results;

size_t result_cache_budget() {
  static const size_t v = [] {
    const char* e = getenv("PICO_B200_RESULT_CACHE_MB");
    return (size_t)(e ? std::max(0L, atol(e)) : 8192L) << 20;
  }();
  return v;
}

void* alloc_result(size_t bytes) {
  constexpr size_t kHuge = (size_t)2 << 20;
  if (bytes >= 32 * kHuge) {
    const size_t cap = (bytes + kHuge - 1) / kHuge * kHuge;
    {
      std::lock_guard<std::mutex> lock(g_results.mu);
      if (g_results.spare && g_results.spare_cap >= cap && g_results.spare_cap / 2 <= cap) {
        void* p = g_results.spare;
        g_results.spare = nullptr;
        g_results.spare_cap = 0;
        return p;  // (still listed in `big`)
      }
    }
    void* p = nullptr;
    if (posix_memalign(&p, kHuge, cap) == 0) {
      madvise(p, cap, MADV_HUGEPAGE);
      std::lock_guard<std::mutex> lock(g_results.mu);
      g_results.big.emplace_back(p, cap);
      return p;
    }
  }
  return malloc(bytes);
}

// Touches every page of a freshly allocated (never written) pageable output buffer with several host
// threads: a device-to-host copy into untouched pages otherwise takes its page faults one at a time
// inside the driver (np.empty results: 160 MB took 200 ms).
void parallel_touch(char* p, size_t bytes) {
  const unsigned nt = copy_threads();
  const size_t per = ((bytes / nt) + 4095) & ~(size_t)4095;
  std::vector<std::thread> th;
  for (unsigned i = 0; i < nt; ++i) {
    const size_t b = std::min(bytes, i * per), e = std::min(bytes, (i + 1) * per);
    if (b < e)
      th.emplace_back([=] {
        for (size_t o = b; o < e; o += 4096) reinterpret_cast<volatile char*>(p)[o] = 0;
      });
  }
  for (auto& t : th) t.join();
}

// Is `p` ordinary pageable host memory (neither pinned nor registered nor managed)?
bool is_pageable(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return true;
  }
  return a.type == cudaMemoryTypeUnregistered;
}

// Pinned mirrors of pageable query / result buffers, grown on demand and kept per host thread:
// cudaMemcpyAsync from pageable memory is staged by the driver on the calling thread at a few
// GB/s and serialises the chunk pipeline; copying with several host threads into pinned memory
// that the copy engines then read at PCIe speed is 2-3x faster end to end for std::vector / numpy
// callers.
struct PinnedMirror {
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    PICO_CUDA(cudaHostAlloc(&p, bytes, cudaHostAllocDefault));
    cap = bytes;
    return 0;
  }
};
thread_local PinnedMirror g_pin_in, g_pin_out;
// Pinning costs close to a millisecond per megabyte, once. The mirrors are therefore bounded (1 GiB), and a
// host thread only gets them from its second big pageable batch on: a one-off call takes the plain
// pageable copies, a caller that keeps searching pays the allocation once and is 1.5x faster after.
constexpr size_t kMirrorBudget = (size_t)1 << 30;
thread_local int g_pageable_batches = 0;

// Small batches from host memory (the reference's one-query-per-call loops end up here): the
// queries are written into a pinned, device-mapped buffer that the kernel reads and writes over
// PCIe directly — no allocation, no copy calls, no events; one launch and one synchronisation.
constexpr size_t kSmallQueryBytes = 16 * 1024, kSmallResultBytes = 48 * 1024;
struct SmallBuffer {
  void* p = nullptr;  // kSmallQueryBytes of queries, then kSmallResultBytes of results
  int ensure() {
    if (!p) PICO_CUDA(cudaHostAlloc(&p, kSmallQueryBytes + kSmallResultBytes, cudaHostAllocMapped));
    return 0;
  }
};
thread_local SmallBuffer g_small;

int copy_out(cudaStream_t st, void* h_dst, const void* d_src, size_t bytes) {
  if (bytes <= ((size_t)8 << 20)) {
    PICO_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, st));
    PICO_CUDA(cudaStreamSynchronize(st));
    return 0;
  }
  std::lock_guard<std::mutex> lock(g_staging.mu);
  for (auto& b : g_staging.buf)
    if (!b) PICO_CUDA(cudaHostAlloc(&b, kStageBytes, cudaHostAllocDefault));
  // The device fills the two staging buffers in turn (stage i -> buffer i & 1); a team of host threads that lives
  // for the whole call moves 2 MiB blocks of completed stages to their destination; the buffer of stage i is handed
  // to stage i + 2 once all of its blocks are out. (A team spawned and joined per stage: 96 x 16 thread starts for
  // the 6.1 GB of cfg3's radius result.)
  cudaEvent_t ev[2];
  PICO_CUDA(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
  PICO_CUDA(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
  char* dst = static_cast<char*>(h_dst);
  const char* src = static_cast<const char*>(d_src);
  constexpr size_t kBlock = (size_t)2 << 20, kBlocksPerStage = kStageBytes / kBlock;
  const size_t n_stages = (bytes + kStageBytes - 1) / kStageBytes, n_blocks = (bytes + kBlock - 1) / kBlock;
  std::vector<std::atomic<uint32_t>> left(n_stages);
  std::vector<std::atomic<uint8_t>> landed(n_stages);  // 1: the stage's device-to-host copy has completed
  for (size_t i = 0; i < n_stages; ++i) {
    left[i].store((uint32_t)std::min(kBlocksPerStage, n_blocks - i * kBlocksPerStage), std::memory_order_relaxed);
    landed[i].store(0, std::memory_order_relaxed);
  }
  std::atomic<size_t> issued{0}, next{0};
  std::atomic<bool> failed{false};
  auto stage_ready = [&](size_t i) -> bool {
    if (landed[i].load(std::memory_order_acquire)) return true;
    if (issued.load(std::memory_order_acquire) <= i) return false;
    // (ev[i & 1] is recorded again for stage i + 2 only after every block of stage i is out)
    const cudaError_t q = cudaEventQuery(ev[i & 1]);
    if (q == cudaSuccess) {
      landed[i].store(1, std::memory_order_release);
      return true;
    }
    if (q != cudaErrorNotReady) failed.store(true);
    cudaGetLastError();
    return false;
  };
  // 1: moved a block; 0: the next block's stage has not landed; -1: nothing left
  auto move_one = [&]() -> int {
    const size_t peek = next.load(std::memory_order_relaxed);
    if (peek >= n_blocks) return -1;
    if (!stage_ready(peek / kBlocksPerStage)) return 0;
    const size_t b = next.fetch_add(1, std::memory_order_relaxed);
    if (b >= n_blocks) return -1;
    const size_t i = b / kBlocksPerStage;
    while (!stage_ready(i)) {
      if (failed.load()) return -1;
      std::this_thread::yield();
    }
    const size_t off = b * kBlock;
    memcpy(dst + off, static_cast<const char*>(g_staging.buf[i & 1]) + (off - i * kStageBytes), std::min(kBlock, bytes - off));
    left[i].fetch_sub(1, std::memory_order_release);
    return 1;
  };
  int device = 0;
  cudaGetDevice(&device);
  std::vector<std::thread> team;
  for (unsigned w = 1; w < result_copy_threads(); ++w)
    team.emplace_back([&, device] {
      cudaSetDevice(device);
      for (;;) {
        if (failed.load(std::memory_order_relaxed)) return;
        const int r = move_one();
        if (r < 0) return;
        if (r == 0) std::this_thread::yield();
      }
    });
  int rc = 0;
  for (size_t i = 0; i < n_stages && !rc; ++i) {
    // (this thread does not move blocks while stages remain to be issued: a block it drew could belong to a stage
    // that only it can issue)
    while (i >= 2 && left[i - 2].load(std::memory_order_acquire) != 0 && !failed.load()) std::this_thread::yield();
    const size_t off = i * kStageBytes, len = std::min(kStageBytes, bytes - off);
    if (failed.load() || cudaMemcpyAsync(g_staging.buf[i & 1], src + off, len, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaEventRecord(ev[i & 1], st) != cudaSuccess) {
      rc = fail(PICO_B200_ERR_CUDA, "staged device-to-host copy failed");
      failed.store(true);
    }
    issued.store(i + 1, std::memory_order_release);
  }
  while (!rc && !failed.load()) {
    const int r = move_one();
    if (r < 0) break;
    if (r == 0) std::this_thread::yield();
  }
  for (auto& th : team) th.join();
  if (!rc && failed.load()) rc = fail(PICO_B200_ERR_CUDA, "staged copy sync failed");
  cudaStreamSynchronize(st);
  cudaEventDestroy(ev[0]);
  cudaEventDestroy(ev[1]);
  return rc;
}

template <typename T>
int stage_queries(CallCtx& c, const T* q, size_t nq, size_t stride, size_t sdim, bool on_device, const T** d_q,
                  size_t* d_stride) {
  if (on_device) {
    *d_q = q;
    *d_stride = stride;
    return 0;
  }
  T* buf = nullptr;
  PICO_TRY(c.alloc(reinterpret_cast<void**>(&buf), nq * sdim * sizeof(T)));
  if (nq && stride == sdim)
    PICO_CUDA(cudaMemcpyAsync(buf, q, nq * sdim * sizeof(T), cudaMemcpyHostToDevice, c.st));
  else if (nq)
    PICO_CUDA(cudaMemcpy2DAsync(buf, sdim * sizeof(T), q, stride * sizeof(T), sdim * sizeof(T), nq,
                                cudaMemcpyHostToDevice, c.st));
  *d_q = buf;
  *d_stride = sdim;
  return 0;
}

// The batch is sorted on the top bits of the 30-bit code (stable radix sort, 8 bits per pass): 16 bits
// (2 passes) for single-neighbour searches, 24 bits (3 passes) for everything else. At k = 1 the third pass
// costs a resident batch what the finer order saves in the traversal (1.50 vs 1.51 ms per 7.2M queries),
// and the 1 Mi chunks of the host pipeline, where the small sort kernels are latency-bound, gain 5 %
// (profiles/r1/host_pipeline_sweep.txt); the heavier traversals (k = 16: 6.3 ms, radius) keep the finer
// order. A counting sort into 2^16 cells (warp-aggregated atomics, one-block scan, scatter) was slower
// than CUB's two passes (1.76 ms resident, host_pipeline_sweep2.txt) and was dropped.
// PICO_B200_MORTON_BITS overrides both (tuning hook).
constexpr int kMortonBitsNn = 16, kMortonBits = 24;
int morton_bits(bool single_neighbour) {
  static const int forced = [] {
    const char* e = getenv("PICO_B200_MORTON_BITS");
    const int x = e ? atoi(e) : 0;
    return (x >= 3 && x <= 30) ? x : 0;
  }();
  return forced ? forced : (single_neighbour ? kMortonBitsNn : kMortonBits);
}

// Z-order permutation of a batch (device). Workspace: three key / value scratch arrays + CUB's temporary storage
// (`scratch`) and the array that receives the permutation (`perm_out`, nq entries).
struct PermPlan {
  size_t arr = 0, tmp_bytes = 0;
  size_t scratch_bytes() const { return 3 * arr + tmp_bytes; }
};
inline int plan_perm(size_t nq, int bits, PermPlan* plan) {
  size_t tmp_bytes = 0;
  PICO_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (uint32_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr,
                                            (uint32_t*)nullptr, (int)nq, 30 - bits, 30, (cudaStream_t) nullptr));
  plan->arr = (nq * 4 + 255) & ~(size_t)255;
  plan->tmp_bytes = tmp_bytes;
  return 0;
}
template <typename T>
int enqueue_perm(cudaStream_t st, const pico_b200_tree* t, const T* d_q, size_t stride, size_t nq, int bits,
                 const PermPlan& plan, char* scratch, uint32_t* perm_out) {
  const int dims = (int)std::min<size_t>(t->sdim, 3);
  T lo[3] = {0, 0, 0}, inv[3] = {0, 0, 0};
  for (int j = 0; j < dims; ++j) {
    lo[j] = (T)t->root_box_host[j];
    const double ext = t->root_box_host[4 + j] - t->root_box_host[j];
    inv[j] = ext > 0 ? (T)(1023.999 / ext) : T(0);
  }
  uint32_t* codes = reinterpret_cast<uint32_t*>(scratch);
  uint32_t* ids = reinterpret_cast<uint32_t*>(scratch + plan.arr);
  uint32_t* codes2 = reinterpret_cast<uint32_t*>(scratch + 2 * plan.arr);
  void* tmp = scratch + 3 * plan.arr;
  size_t tmp_bytes = plan.tmp_bytes;
  morton_kernel<T><<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(d_q, stride, (uint32_t)nq, dims, lo[0], lo[1], lo[2],
                                                                 inv[0], inv[1], inv[2], codes, ids);
  PICO_CUDA(cudaGetLastError());
  PICO_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, codes, codes2, ids, perm_out, (int)nq, 30 - bits, 30, st));
  return 0;
}

// ---- tile-local ordering
// Batches usually arrive in an order that is already coherent at a coarse scale (a LiDAR scan, a raster, the
// output of a previous spatial pass): consecutive queries lie in the same region, only not next to each other.
// Then a global sort is more than is needed. tile_order_kernel sorts every tile of 2048 consecutive queries by its
// Morton code on its own (one block, one cub::BlockRadixSort over 20 bits, no global pass, ONE launch instead of
// the Morton kernel + two device-wide radix passes): on the cfg2 scan-order batch the 32 queries of a warp then
// share 5.7 cells of 0.25 m on average, against 9.0 after the global 16-bit sort and 5.3 after a global 24-bit
// sort (input order: 25.8; profiles/r2/order_quality.txt). It also says how coherent the batch was: the number of
// distinct coarse cells (top 15 code bits) that 32 consecutive sorted ranks — one traversal warp — touch, summed
// into `stat`. A shuffled batch touches ~32 and gains nothing from a local sort; pico_b200_order_hint remembers that and such trees keep the global sort.
constexpr int kTileItems = 8, kTileThreads = 256, kTile = kTileItems * kTileThreads;

template <typename T>
__global__ void __launch_bounds__(kTileThreads) tile_order_kernel(const T* __restrict__ q, size_t stride, uint32_t nq,
                                                                 int dims, T lo0, T lo1, T lo2, T inv0, T inv1, T inv2,
                                                                 uint32_t* __restrict__ perm,
                                                                 unsigned long long* __restrict__ stat) {
  using Sort = cub::BlockRadixSort<uint32_t, kTileThreads, kTileItems, uint16_t>;  // (6-bit digits: 124 registers, no faster)
  __shared__ typename Sort::TempStorage sort_tmp;
  const uint32_t base = blockIdx.x * (uint32_t)kTile;
  const T l[3] = {lo0, lo1, lo2}, s[3] = {inv0, inv1, inv2};
  uint32_t keys[kTileItems];
  uint16_t vals[kTileItems];
#pragma unroll
  for (int i = 0; i < kTileItems; ++i) {
    const uint32_t local = (uint32_t)i * kTileThreads + threadIdx.x;  // striped: coalesced loads
    uint32_t code = 0xFFFFFFFFu;                                       // past the end: sorts last
    if (base + local < nq) {
      const T* p = q + (size_t)(base + local) * stride;
      code = 0;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        if (j < dims) {
          T f = (p[j] - l[j]) * s[j];
          f = f < T(0) ? T(0) : (f > T(1023) ? T(1023) : f);
          code |= spread10((uint32_t)f) << j;
        }
      }
    }
    keys[i] = code;
    vals[i] = (uint16_t)local;
  }
  // (the arrangement going in does not matter to a sort.) 20 bits = cells of ~1/128 of the extent per dimension are
  // as fine as a warp of 32 queries can use: five 4-bit passes. Only a partial last tile needs the two top bits,
  // which keep its padding behind the valid entries.
  if (base + kTile <= nq)
    Sort(sort_tmp).SortBlockedToStriped(keys, vals, 10, 30);
  else
    Sort(sort_tmp).SortBlockedToStriped(keys, vals, 10, 32);
  // thread t now holds ranks t, t + 128, ...: the 32 lanes of a warp hold 32 consecutive ranks for every i —
  // exactly the groups of queries a traversal warp will work on. Count the coarse cells such a group touches.
  uint32_t cells = 0;
#pragma unroll
  for (int i = 0; i < kTileItems; ++i) {
    const uint32_t rank = (uint32_t)i * kTileThreads + threadIdx.x;
    if (base + rank < nq) perm[base + rank] = base + vals[i];  // valid entries sort in front of the padding
    const uint32_t c = keys[i] >> 15, prev = __shfl_up_sync(0xffffffffu, c, 1);
    cells += __popc(__ballot_sync(0xffffffffu, (threadIdx.x & 31) == 0 || c != prev));
  }
  // one word: groups in the top 24 bits, cells below (several streams may measure into the same word; a single
  // word cannot be seen half-updated)
  if ((threadIdx.x & 31) == 0) atomicAdd(stat, ((unsigned long long)kTileItems << 40) + (unsigned long long)cells);
}

// ---- exact nn with the far subtrees as work items of the block (float, sdim 2 / 3, metric_l2_squared)
// The thread-per-query traversal spends two thirds of its warp instructions below far children at ~3 of 32 lanes:
// after the coherent part (first descent, first leaf, second walk — 29 lanes) every lane is left with 0..4 far
// subtrees of its own, and a warp runs until its slowest lane is done. Here the coherent part only EMITS the far
// children that can matter ({owner thread, node | split_dim << 30, offset}: on the first path the box distance is the
// offset itself) into the block's slice of a global item array; after a barrier the 128 threads of the block take the
// items one each, whoever emitted them, walk the subtree in the reference's order and fold what they find into the
// owner's result with a 64-bit atomicMin on {index, distance} (distance in the high word: non-negative floats order
// like unsigned integers). The subtrees of one query are then searched side by side against the best of the first
// leaf instead of one after the other against a shrinking bound (more nodes visited, at four times the lanes), and
// "first visited wins" is no longer decided by the traversal: a query whose best distance is attained twice — inside
// one item (VisitNnTie), or across items (the atomicMin returns what it replaced) — is flagged in `redo` and re-run
// by the order-exact kernel. So is a query of a block whose item slice overflowed.
constexpr int kItemsPerBlock = 4 * kThreadsPerBlock;

template <int DIM>
__global__ void __launch_bounds__(kThreadsPerBlock, 2048 / kThreadsPerBlock) nn_split_kernel(KnnArgs<float> a,
                                                                                             uint32_t* __restrict__ items,
                                                                                             uint8_t* __restrict__ redo) {
  using T = float;
  constexpr int metric = PICO_B200_METRIC_L2_SQUARED;
  __shared__ uint32_t n_items;
  if (threadIdx.x == 0) n_items = 0;
  __syncthreads();
  uint32_t* my_items = items + (size_t)blockIdx.x * kItemsPerBlock * 3;
  const uint32_t slot = blockIdx.x * kThreadsPerBlock + threadIdx.x;
  if (slot < a.nq) {
    const uint32_t qi = a.perm ? a.perm[slot] : slot;
    T q[DIM];
    const T* qp = a.q + (size_t)qi * a.q_stride;
#pragma unroll
    for (int j = 0; j < DIM; ++j) q[j] = qp[j];
    // first descent, first leaf
    T na, nb;
    uint32_t right, sd, node = 0;
    int lb, le;
    load_node(a.nodes, node, na, nb, right, sd, lb, le);
    while (sd != PICO_B200_LEAF) {
      T v = q[0];
#pragma unroll
      for (int j = 1; j < DIM; ++j) v = (sd == (uint32_t)j) ? q[j] : v;
      node = sub_rn(sub_rn(add_rn(na, nb), v), v) > T(0) ? node + 1 : right;
      load_node(a.nodes, node, na, nb, right, sd, lb, le);
    }
    VisitNnTie<T> vis;
    for (int i = lb; i < le; ++i) {
      const float4 p = ldg4(a.pts4 + i);
      T d = metric_first(metric, q[0], p.x);
      if (DIM > 1) d = metric_fold(metric, d, q[DIM > 1 ? 1 : 0], p.y, 1);
      if (DIM > 2) d = metric_fold(metric, d, q[DIM > 2 ? 2 : 0], p.z, 2);
      vis.visit(index_of(p), d);
    }
    const T reach = add_rn(vis.best, mul_rn(vis.best, T(1.2207031e-4)));
    // second walk: far children of the first path within reach become items (deepest last; the order is free now)
    bool overflow = false;
    node = 0;
    load_node(a.nodes, node, na, nb, right, sd, lb, le);
    while (sd != PICO_B200_LEAF) {
      T v = q[0];
#pragma unroll
      for (int j = 1; j < DIM; ++j) v = (sd == (uint32_t)j) ? q[j] : v;
      const bool go_left = sub_rn(sub_rn(add_rn(na, nb), v), v) > T(0);
      const T t = sub_rn(go_left ? nb : na, v);
      const T new_off = mul_rn(t, t);
      if (reach >= new_off) {
        const uint32_t pos = atomicAdd(&n_items, 1u);
        if (pos < (uint32_t)kItemsPerBlock) {
          my_items[3 * pos] = threadIdx.x;
          my_items[3 * pos + 1] = (go_left ? right : node + 1) | (sd << 30);
          my_items[3 * pos + 2] = __float_as_uint(new_off);
        } else {
          overflow = true;
        }
      }
      node = go_left ? node + 1 : right;
      load_node(a.nodes, node, na, nb, right, sd, lb, le);
    }
    Neighbor<T>* out = a.out + qi;
    store_neighbor(out, vis.idx, vis.best);
    if (vis.tie || overflow) redo[qi] = 1;
  }
  __syncthreads();
  const uint32_t n = min(n_items, (uint32_t)kItemsPerBlock);
  for (uint32_t j = threadIdx.x; j < n; j += kThreadsPerBlock) {
    const uint32_t owner = my_items[3 * j], tag = my_items[3 * j + 1];
    const T new_off = __uint_as_float(my_items[3 * j + 2]);
    const uint32_t oslot = blockIdx.x * kThreadsPerBlock + owner;
    const uint32_t qi = a.perm ? a.perm[oslot] : oslot;
    T q[DIM];
    const T* qp = a.q + (size_t)qi * a.q_stride;
#pragma unroll
    for (int jj = 0; jj < DIM; ++jj) q[jj] = qp[jj];
    unsigned long long* cell = reinterpret_cast<unsigned long long*>(a.out + qi);
    const unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(cell);
    VisitNnTie<T> vis;
    vis.best = __uint_as_float((uint32_t)(cur >> 32));
    vis.idx = (int)(uint32_t)cur;
    const T start_best = vis.best;
    T off[DIM];
    const uint32_t sd = tag >> 30;
#pragma unroll
    for (int jj = 0; jj < DIM; ++jj) off[jj] = (sd == (uint32_t)jj) ? new_off : T(0);
    LocalStack<T, DIM, kLocalStack> st;
    traverse_subtree<T, DIM>(a.nodes, a.pts4, q, tag & kTagNodeMask, new_off, off, mul_rn(vis.best, T(1.2207031e-4)), st,
                             vis);
    bool tie = vis.tie;
    if (vis.best < start_best) {
      const unsigned long long mine = ((unsigned long long)__float_as_uint(vis.best) << 32) | (uint32_t)vis.idx;
      const unsigned long long old = atomicMin(cell, mine);
      tie = tie || ((uint32_t)(old >> 32) == __float_as_uint(vis.best) && (uint32_t)old != (uint32_t)vis.idx);
    }
    if (tie) redo[qi] = 1;
  }
}

// The flagged queries of nn_split_kernel, once more through the order-exact traversal.
template <int DIM>
__global__ void __launch_bounds__(kThreadsPerBlock) nn_redo_kernel(KnnArgs<float> a, const uint8_t* __restrict__ redo) {
  using T = float;
  const uint32_t qi = blockIdx.x * kThreadsPerBlock + threadIdx.x;
  if (qi >= a.nq || !redo[qi]) return;
  T q[DIM];
  const T* qp = a.q + (size_t)qi * a.q_stride;
#pragma unroll
  for (int j = 0; j < DIM; ++j) q[j] = qp[j];
  VisitNn<T> vis;
  LocalStack<T, DIM, kLocalStack> st;
  traverse_packed<T, DIM, true, kPrimeFirstLeaf>(a.nodes, a.pts4, a.outer, q, a.metric, false, a.e_inv, st, vis);
  store_neighbor(a.out + qi, vis.idx, vis.best);
}

// ---- order and traverse in one kernel (k = 1, metric_l2_squared, batches that arrive locally coherent)
// Each block takes a tile of kFusedTile consecutive queries, Z-orders it in shared memory (keys = 20 code bits with the
// local index packed below them: a keys-only block radix sort) and then walks the tree for its queries, every warp
// over 32 consecutive ranks. No permutation array, no ordering kernel in front: in the host pipeline a chunk's
// traversal no longer waits for a chain of small sort kernels that cannot find room among the traversal blocks of the
// chunks before it (profiles/r2/host_pipeline_timeline.txt). The tile is smaller than tile_order_kernel's (512 against
// 2048: a warp then touches ~9.0 cells of 0.25 m, what the device-wide 16-bit order gives, profiles/r2/order_quality.txt)
// because every thread walks kFusedItems queries one after the other.
constexpr int kFusedItems = 4, kFusedTile = kFusedItems * kThreadsPerBlock;

template <typename T, int DIM>
__global__ void __launch_bounds__(kThreadsPerBlock, 2048 / kThreadsPerBlock) nn_tile_kernel(KnnArgs<T> a, T lo0, T lo1, T lo2,
                                                                                            T inv0, T inv1, T inv2) {
  using Sort = cub::BlockRadixSort<uint32_t, kThreadsPerBlock, kFusedItems>;
  __shared__ union {
    typename Sort::TempStorage sort;
    uint16_t order[kFusedTile];
  } sm;
  const uint32_t base = blockIdx.x * (uint32_t)kFusedTile;
  const T l[3] = {lo0, lo1, lo2}, s[3] = {inv0, inv1, inv2};
  uint32_t keys[kFusedItems];
#pragma unroll
  for (int i = 0; i < kFusedItems; ++i) {
    const uint32_t local = (uint32_t)i * kThreadsPerBlock + threadIdx.x;
    uint32_t key = 0xFFFFF000u | local;  // past the end of the batch: behind every valid entry
    if (base + local < a.nq) {
      const T* p = a.q + (size_t)(base + local) * a.q_stride;
      uint32_t code = 0;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        if (j < DIM) {
          T f = (p[j] - l[j]) * s[j];
          f = f < T(0) ? T(0) : (f > T(1023) ? T(1023) : f);
          code |= spread10((uint32_t)f) << j;
        }
      }
      key = ((code >> 10) << 12) | local;
    }
    keys[i] = key;
  }
  Sort(sm.sort).SortBlockedToStriped(keys, 12, 32);
  __syncthreads();  // the sort's storage becomes the order array
#pragma unroll
  for (int i = 0; i < kFusedItems; ++i) sm.order[i * kThreadsPerBlock + threadIdx.x] = (uint16_t)(keys[i] & 0xFFFu);
  __syncthreads();
  for (int i = 0; i < kFusedItems; ++i) {
    const uint32_t qi = base + sm.order[i * kThreadsPerBlock + threadIdx.x];
    if (qi >= a.nq) continue;
    T q[DIM];
    const T* qp = a.q + (size_t)qi * a.q_stride;
#pragma unroll
    for (int j = 0; j < DIM; ++j) q[j] = qp[j];
    VisitNn<T> vis;
    LocalStack<T, DIM, kLocalStack> st;
    traverse_packed<T, DIM, true, kPrimeFirstLeaf>(a.nodes, a.pts4, a.outer, q, a.metric, false, a.e_inv, st, vis);
    Neighbor<T>* out = a.out + qi;
    store_neighbor(out, vis.idx, vis.best);
  }
}

// PICO_B200_ORDER (tuning hook): "auto" (default: per-tree hint), "global", "local"
int order_mode() {
  static const int v = [] {
    const char* e = getenv("PICO_B200_ORDER");
    if (e && !strcmp(e, "global")) return 1;
    if (e && !strcmp(e, "local")) return 2;
    return 0;
  }();
  return v;
}

constexpr unsigned long long kCoherentCellsPerGroup = 8;  // coarse cells per 32 consecutive ranks: a scan ~1-3, shuffled ~32

// Reads what the last measured batch looked like and decides for this call: true = tile-local order.
bool order_locally(const pico_b200_tree* t) {
  pico_b200_order_hint& h = t->order_hint;
  const int mode = order_mode();
  if (mode) return mode == 2;
  if (h.h_stat) {
    const unsigned long long word = reinterpret_cast<volatile unsigned long long*>(h.h_stat)[0];
    const unsigned long long sum = word & ((1ull << 40) - 1), tiles = word >> 40;
    if (tiles > 0) h.state.store(sum <= tiles * kCoherentCellsPerGroup ? 1 : 2, std::memory_order_relaxed);
  }
  return h.state.load(std::memory_order_relaxed) == 1;
}

template <typename T>
int enqueue_tile_order(cudaStream_t st, const pico_b200_tree* t, const T* d_q, size_t stride, size_t nq,
                       uint32_t* perm_out) {
  pico_b200_order_hint& h = t->order_hint;
  if (!h.d_stat) {
    // (two host threads may both get here: one of the two small allocations is then leaked, nothing worse)
    unsigned long long *d = nullptr, *hp = nullptr;
    PICO_CUDA(cudaMalloc(reinterpret_cast<void**>(&d), sizeof(unsigned long long)));
    PICO_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&hp), sizeof(unsigned long long), cudaHostAllocDefault));
    hp[0] = 0;
    h.h_stat = hp;
    h.d_stat = d;
  }
  const int dims = (int)std::min<size_t>(t->sdim, 3);
  T lo[3] = {0, 0, 0}, inv[3] = {0, 0, 0};
  for (int j = 0; j < dims; ++j) {
    lo[j] = (T)t->root_box_host[j];
    const double ext = t->root_box_host[4 + j] - t->root_box_host[j];
    inv[j] = ext > 0 ? (T)(1023.999 / ext) : T(0);
  }
  PICO_CUDA(cudaMemsetAsync(h.d_stat, 0, sizeof(unsigned long long), st));
  tile_order_kernel<T><<<(unsigned)((nq + kTile - 1) / kTile), kTileThreads, 0, st>>>(
      d_q, stride, (uint32_t)nq, dims, lo[0], lo[1], lo[2], inv[0], inv[1], inv[2], perm_out, h.d_stat);
  PICO_CUDA(cudaGetLastError());
  PICO_CUDA(cudaMemcpyAsync(h.h_stat, h.d_stat, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  return 0;
}

// Returns nullptr in *perm for tiny batches.
template <typename T>
int make_perm(CallCtx& c, const pico_b200_tree* t, const T* d_q, size_t stride, size_t nq, unsigned flags,
              uint32_t** perm, bool single_neighbour = false) {
  const int bits = morton_bits(single_neighbour);
  *perm = nullptr;
  if ((flags & PICO_B200_NO_REORDER) || nq < 2048) return 0;
  // (k > 1 keeps the 24-bit device-wide order: 6.71 against 6.77 ms at k = 16, profiles/r2/order_sweep_v2.txt)
  const bool local = !c.pipelined && single_neighbour && order_locally(t);
  // every 16th call of a tree that keeps the global sort measures the batch again (one extra small kernel)
  const bool probe = !local && !c.pipelined && single_neighbour && order_mode() == 0 && (t->order_hint.calls.fetch_add(1, std::memory_order_relaxed) % 16 == 0);
  PermPlan plan;
  if (!local) PICO_TRY(plan_perm(nq, bits, &plan));
  const size_t arr = (nq * 4 + 255) & ~(size_t)255;
  char* ws = nullptr;
  // layout: the permutation, [the probe's permutation,] the sort's scratch arrays, CUB's temporary storage last
  // (it has no particular size, nothing may be placed behind it)
  PICO_TRY(c.alloc(reinterpret_cast<void**>(&ws), arr + (probe ? arr : 0) + (local ? 0 : plan.scratch_bytes())));
  uint32_t* out = reinterpret_cast<uint32_t*>(ws);
  if (local) {
    PICO_TRY(enqueue_tile_order<T>(c.st, t, d_q, stride, nq, out));
  } else {
    PICO_TRY(enqueue_perm<T>(c.st, t, d_q, stride, nq, bits, plan, ws + arr + (probe ? arr : 0), out));
    if (probe) PICO_TRY(enqueue_tile_order<T>(c.st, t, d_q, stride, nq, reinterpret_cast<uint32_t*>(ws + arr)));
  }
  *perm = out;
  return 0;
}

template <typename T>
void fill_base(KnnArgs<T>& a, const pico_b200_tree* t, const T* d_q, size_t d_stride, size_t nq, const uint32_t* perm,
               double e) {
  a.nodes = static_cast<const typename NodeOf<T>::type*>(t->d_nodes);
  a.outer = static_cast<const T*>(t->d_outer);
  a.spans = nullptr;
  a.pts4 = t->packed() ? static_cast<const typename Vec4Of<T>::type*>(t->d_pts) : nullptr;
  a.rows = t->packed() ? nullptr : static_cast<const T*>(t->d_pts);
  a.indices = t->d_indices;
  a.q = d_q;
  a.q_stride = d_stride;
  a.nq = (uint32_t)nq;
  a.perm = perm;
  a.out = nullptr;
  a.k = 0;
  a.sdim = (int)t->sdim;
  a.n_points = (uint32_t)t->n;
  a.metric = t->metric;
  a.approx = e > 0;
  a.e_inv = e > 0 ? T(1.0) / T(e) : T(1.0);  // search_visitor.hpp:173,216,265
  a.ws = nullptr;
  a.ws_stride = a.ws_depth = 0;
  a.warp_smem = 2 * (int)t->sdim + (int)(kStackWindow * sizeof(WarpFrame<T>) / sizeof(T));
  a.tile_rows = 0;
  a.cache_rows = 0;
  a.counter = nullptr;
  a.fat = a.far_nodes = nullptr;
  a.tie_count = a.tie_list = nullptr;
  a.nq_from = nullptr;
}

// Shared memory per warp of the warp-per-query kernels, in bytes; decides whether leaves are
// staged through shared memory (row storage, 16-byte rows, the tile fits next to 8 warps' state).
template <typename T>
size_t plan_warp_smem(const pico_b200_tree* t, KnnArgs<T>& a) {
  const size_t sdim = t->sdim;
  const size_t vec = 16 / sizeof(T);
  const size_t window = kStackWindow * sizeof(WarpFrame<T>) / sizeof(T);  // scalars; a multiple of 16 bytes
  a.warp_smem = (int)(2 * sdim + window);
  a.tile_rows = 0;
  a.cache_rows = 0;
  a.spans = nullptr;
  if (!t->packed() && sdim % vec == 0 && t->max_leaf_points > 0) {
    // per warp: query + offsets, the tile (rows x (sdim + pad)), the distance / index cache of the tile's
    // rows, the stack window. The tile holds a whole subtree when it can (subtree distance cache): as many
    // rows as fit while FOUR blocks of 8 warps stay resident per SM (the kernel's 64 registers allow no
    // more; a bigger tile costs occupancy and was slower, profiles/r1/hd_tile_rows.txt), at most 32.
    static const size_t cap = [] {
      const char* e = getenv("PICO_B200_TILE_ROWS");  // tuning hook
      const int x = e ? atoi(e) : 0;
      return (size_t)((x >= 1 && x <= 32) ? x : 32);
    }();
    // distances of up to cache_factor tiles are kept: in high dimensions nothing gets pruned, and a span
    // of several small leaves then costs one full-tile round per tile_rows points instead of one per leaf
    static const size_t cache_factor = [] {
      const char* e = getenv("PICO_B200_CACHE_TILES");  // tuning hook
      const int x = e ? atoi(e) : 0;
      return (size_t)((x >= 1 && x <= 64) ? x : 16);  // profiles/r1/hd_tile_rows.txt
    }();
    const size_t factor = sdim >= 64 ? cache_factor : 1;
    auto slice = [&](size_t rows) {
      const size_t cache = (2 * rows * factor + vec - 1) / vec * vec;
      return 2 * sdim + rows * (sdim + vec) + cache + window;
    };
    size_t rows = cap;
    constexpr size_t kBlockBudget = 56 * 1024;
    while (rows > 1 && slice(rows) * sizeof(T) * kWarpsPerBlock > kBlockBudget) --rows;
    if (slice(rows) * sizeof(T) * kWarpsPerBlock <= kBlockBudget) {
      a.tile_rows = (int)rows;
      a.cache_rows = (int)(rows * factor);
      a.warp_smem = (int)slice(rows);
      a.spans = t->d_spans;
    }
  }
  return (size_t)a.warp_smem * sizeof(T);
}

// thread-per-query launch geometry + optional deep-tree workspace
template <typename T>
int thread_geometry(CallCtx& c, const pico_b200_tree* t, KnnArgs<T>& a, int dim, bool* deep, unsigned* blocks) {
  *deep = t->height >= (size_t)kLocalStack;
  size_t threads = ((size_t)a.nq + kThreadsPerBlock - 1) / kThreadsPerBlock * kThreadsPerBlock;
  if (*deep) {
    const size_t depth = t->height + 1;
    const size_t per_thread = depth * (4 + sizeof(T) * (1 + dim));
    const size_t budget = (size_t)2 << 30;
    size_t max_threads = std::max<size_t>(budget / per_thread / kThreadsPerBlock, 1) * kThreadsPerBlock;
    max_threads = std::min<size_t>(max_threads, (size_t)t->sm_count * 2048);
    threads = std::min(threads, max_threads);
    a.ws_stride = threads;
    a.ws_depth = depth;
    PICO_TRY(c.alloc(&a.ws, threads * per_thread));
  }
  *blocks = (unsigned)std::max<size_t>(threads / kThreadsPerBlock, 1);
  return 0;
}

template <typename T>
int warp_geometry(CallCtx& c, const pico_b200_tree* t, size_t items, size_t frame_bytes, size_t smem_per_warp,
                  void** ws, size_t* depth, unsigned* blocks, size_t* smem) {
  *depth = t->height + 2;
  size_t warps = std::min<size_t>((items + 0), (size_t)t->sm_count * 64);
  size_t nblocks = std::max<size_t>((warps + kWarpsPerBlock - 1) / kWarpsPerBlock, 1);
  *smem = smem_per_warp * kWarpsPerBlock;
  if (*smem > 200 * 1024) return fail(PICO_B200_ERR_UNSUPPORTED, "spatial dimension too large for shared memory");
  *blocks = (unsigned)nblocks;
  PICO_TRY(c.alloc(ws, nblocks * kWarpsPerBlock * *depth * frame_bytes));
  return 0;
}

// Persistent launch: no more blocks than are resident at once (occupancy x SMs); the warps draw
// queries from `counter`, which is zeroed here.
template <typename Kernel>
int persistent_grid(CallCtx& c, const pico_b200_tree* t, Kernel kernel, size_t smem, unsigned long long** counter,
                    unsigned* blocks) {
  if (smem > 48 * 1024)
    PICO_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  PICO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kWarpsPerBlock * 32, smem));
  if (per_sm < 1) return fail(PICO_B200_ERR_UNSUPPORTED, "warp kernel does not fit on an SM");
  *blocks = std::min<unsigned>(*blocks, (unsigned)per_sm * (unsigned)t->sm_count);
  if (!*counter) PICO_TRY(c.alloc(reinterpret_cast<void**>(counter), sizeof(unsigned long long)));
  PICO_CUDA(cudaMemsetAsync(*counter, 0, sizeof(unsigned long long), c.st));
  return 0;
}

template <typename T, int DIM, bool FAST, bool DEEP>
void launch_knn_thread_k(const KnnArgs<T>& a, unsigned blocks, cudaStream_t st) {
  if (a.k == 1)
    knn_thread_kernel<T, DIM, 1, FAST, DEEP><<<blocks, kThreadsPerBlock, 0, st>>>(a);
  else if (a.k <= 4)
    knn_thread_kernel<T, DIM, 4, FAST, DEEP><<<blocks, kThreadsPerBlock, 0, st>>>(a);
  else if (a.k <= 8)
    knn_thread_kernel<T, DIM, 8, FAST, DEEP><<<blocks, kThreadsPerBlock, 0, st>>>(a);
  else if (a.k <= 16)
    knn_thread_kernel<T, DIM, 16, FAST, DEEP><<<blocks, kThreadsPerBlock, 0, st>>>(a);
  else
    knn_thread_kernel<T, DIM, 32, FAST, DEEP><<<blocks, kThreadsPerBlock, 0, st>>>(a);
}

template <typename T, int DIM>
void launch_knn_thread(const KnnArgs<T>& a, bool fast, bool deep, unsigned blocks, cudaStream_t st) {
  if (fast) {
    if (deep)
      launch_knn_thread_k<T, DIM, true, true>(a, blocks, st);
    else
      launch_knn_thread_k<T, DIM, true, false>(a, blocks, st);
  } else {
    if (deep)
      launch_knn_thread_k<T, DIM, false, true>(a, blocks, st);
    else
      launch_knn_thread_k<T, DIM, false, false>(a, blocks, st);
  }
}

// PICO_B200_NN (tuning hook, default 0): 0 = knn_thread_kernel (order-exact, per-thread local stack) — the fastest
// (profiles/r2/nn_sweep_*.txt); bit 0 = nn_kernel (three-word slot stack in shared memory, restore records);
// bit 1 = far children are walked in the search image too (default: in the real tree); bit 2 = no prefix-minimum
// restart records; bit 3 = ignore the search image even if the tree has one (PICO_B200_FAT_LEAF);
// bit 4 = let nn_kernel use up to 40 registers (12 resident blocks per SM instead of 16);
// bit 5 = nn_split_kernel (far subtrees as work items of the block, float only)
int nn_mode() {
  static const int v = [] {
    const char* e = getenv("PICO_B200_NN");
    const int x = e ? atoi(e) : -1;
    return (x >= 0 && x <= 63) ? x : 0;
  }();
  return v;
}

// nn_split_kernel + nn_redo_kernel (PICO_B200_NN bit 5). Returns false if the call is not of that kind.
template <typename T>
bool launch_nn_split(CallCtx&, const pico_b200_tree*, KnnArgs<T>&, bool, bool, size_t, int, uint64_t*) {
  return false;
}
template <>
bool launch_nn_split<float>(CallCtx& c, const pico_b200_tree* t, KnnArgs<float>& a, bool fast, bool deep, size_t k,
                            int mode, uint64_t* launches) {
  if (!(mode & 32) || !fast || deep || k != 1 || t->sdim < 2 || t->n_nodes >= ((size_t)1 << 30) ||
      (reinterpret_cast<uintptr_t>(a.out) & 7u) != 0)
    return false;
  const unsigned blocks = (a.nq + kThreadsPerBlock - 1) / kThreadsPerBlock;
  uint32_t* items = nullptr;
  uint8_t* redo = nullptr;
  c.failed = c.alloc(reinterpret_cast<void**>(&items), (size_t)blocks * kItemsPerBlock * 3 * sizeof(uint32_t));
  if (!c.failed) c.failed = c.alloc(reinterpret_cast<void**>(&redo), a.nq);
  if (!c.failed && cudaMemsetAsync(redo, 0, a.nq, c.st) != cudaSuccess) c.failed = fail(PICO_B200_ERR_CUDA, "memset failed");
  if (c.failed) return true;
  if (t->sdim == 2) {
    nn_split_kernel<2><<<blocks, kThreadsPerBlock, 0, c.st>>>(a, items, redo);
    nn_redo_kernel<2><<<blocks, kThreadsPerBlock, 0, c.st>>>(a, redo);
  } else {
    nn_split_kernel<3><<<blocks, kThreadsPerBlock, 0, c.st>>>(a, items, redo);
    nn_redo_kernel<3><<<blocks, kThreadsPerBlock, 0, c.st>>>(a, redo);
  }
  if (cudaGetLastError() != cudaSuccess) c.failed = fail(PICO_B200_ERR_CUDA, "nn_split_kernel launch failed");
  *launches += 1;
  return true;
}

float elapsed(cudaEvent_t a, cudaEvent_t b) {
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}

// Enqueues one knn batch on c.st: stage queries (host pointers), Z-order, traverse, copy back.
// `perm_in` (with have_perm): the batch was ordered elsewhere (host pipeline: on a high-priority stream).
template <typename T>
int knn_enqueue(CallCtx& c, const pico_b200_tree* t, const T* q, size_t nq, size_t stride, size_t k, double e,
                Neighbor<T>* out, unsigned flags, bool on_device, uint64_t* launches,
                const uint32_t* perm_in = nullptr, bool have_perm = false) {
  PICO_TRY(c.mark(0));
  const T* d_q = nullptr;
  size_t d_stride = 0;
  PICO_TRY(stage_queries(c, q, nq, stride, t->sdim, on_device, &d_q, &d_stride));
  Neighbor<T>* d_out = out;
  if (!on_device) PICO_TRY(c.alloc(reinterpret_cast<void**>(&d_out), nq * k * sizeof(Neighbor<T>)));
  PICO_TRY(c.mark(1));
  if (c.probe_order && nq >= 2048 && t->packed()) {
    uint32_t* scratch_perm = nullptr;
    PICO_TRY(c.alloc(reinterpret_cast<void**>(&scratch_perm), nq * sizeof(uint32_t)));
    PICO_TRY(enqueue_tile_order<T>(c.st, t, d_q, d_stride, nq, scratch_perm));
  }
  uint32_t* perm = const_cast<uint32_t*>(perm_in);
  // order-and-traverse in one kernel (nn_tile_kernel): exact nn, packed float/double points, sdim 2 or 3, a tree the
  // local stack can hold, and a batch the tree expects to be locally coherent
  const bool fused = c.fused_order && !have_perm && k == 1 && !(e > 0) && t->metric == PICO_B200_METRIC_L2_SQUARED &&
                     t->packed() && t->sdim >= 2 && t->height < (size_t)kLocalStack && nq >= 2048 &&
                     !(flags & (PICO_B200_NO_REORDER | PICO_B200_WARP_PER_QUERY)) && nn_mode() == 0;
  if (!have_perm && !fused) PICO_TRY(make_perm(c, t, d_q, d_stride, nq, flags, &perm, k == 1));
  PICO_TRY(c.mark(2));
  PICO_TRY(c.span_begin());
  if (fused) {
    KnnArgs<T> a;
    fill_base(a, t, d_q, d_stride, nq, nullptr, e);
    a.out = d_out;
    a.k = 1;
    T lo[3] = {0, 0, 0}, inv[3] = {0, 0, 0};
    for (int j = 0; j < (int)t->sdim; ++j) {
      lo[j] = (T)t->root_box_host[j];
      const double ext = t->root_box_host[4 + j] - t->root_box_host[j];
      inv[j] = ext > 0 ? (T)(1023.999 / ext) : T(0);
    }
    const unsigned tiles = (unsigned)((nq + kFusedTile - 1) / kFusedTile);
    if (t->sdim == 2)
      nn_tile_kernel<T, 2><<<tiles, kThreadsPerBlock, 0, c.st>>>(a, lo[0], lo[1], lo[2], inv[0], inv[1], inv[2]);
    else
      nn_tile_kernel<T, 3><<<tiles, kThreadsPerBlock, 0, c.st>>>(a, lo[0], lo[1], lo[2], inv[0], inv[1], inv[2]);
    PICO_CUDA(cudaGetLastError());
    PICO_TRY(c.span_end());
    *launches += 1;
    PICO_TRY(c.mark(3));
    if (!on_device)
      PICO_CUDA(cudaMemcpyAsync(out, d_out, nq * k * sizeof(Neighbor<T>), cudaMemcpyDeviceToHost, c.st));
    PICO_TRY(c.mark(4));
    return 0;
  }

  KnnArgs<T> a;
  fill_base(a, t, d_q, d_stride, nq, perm, e);
  a.out = d_out;
  a.k = (int)k;
  *launches += perm ? 3 : 0;
  const bool use_thread = t->packed() && k <= kThreadKMax && !(flags & PICO_B200_WARP_PER_QUERY);
  if (use_thread) {
    bool deep;
    unsigned blocks;
    PICO_TRY(thread_geometry(c, t, a, (int)t->sdim, &deep, &blocks));
    if (c.blocks_per_sm_cap > 0)
      blocks = std::min<unsigned>(blocks, (unsigned)t->sm_count * (unsigned)c.blocks_per_sm_cap);
    const bool fast = t->metric == PICO_B200_METRIC_L2_SQUARED && !(e > 0);
    const int mode = nn_mode();
    if (launch_nn_split(c, t, a, fast, deep, k, mode, launches)) {
      if (c.failed) return c.failed;
    } else if (fast && !deep && k == 1 && (mode & 31) && t->sdim >= 2 && t->n_nodes < ((size_t)1 << 30)) {
      // exact nn with the shared-memory slot stack; with a search image (fat.cu) ties at the best distance go
      // through the order-exact kernel afterwards
      const bool use_fat = t->d_fat_nodes != nullptr && !(mode & 8);
      a.fat = use_fat ? static_cast<const typename NodeOf<T>::type*>(t->d_fat_nodes) : a.nodes;
      a.far_nodes = (mode & 2) ? a.fat : a.nodes;
      if (use_fat) {
        uint32_t* tie = nullptr;
        PICO_TRY(c.alloc(reinterpret_cast<void**>(&tie), (nq + 1) * sizeof(uint32_t)));
        PICO_CUDA(cudaMemsetAsync(tie, 0, sizeof(uint32_t), c.st));
        a.tie_count = tie;
        a.tie_list = tie + 1;
      }
      const bool rec = !(mode & 4);
      // MINB: resident blocks per SM the register allocation must allow (16 = full occupancy, 32 registers)
#define PICO_LAUNCH_NN2(D, MINB)                                                           \
  do {                                                                                     \
    if (use_fat) {                                                                         \
      if (rec)                                                                             \
        nn_kernel<T, D, 3, true, MINB><<<blocks, kThreadsPerBlock, 0, c.st>>>(a);          \
      else                                                                                 \
        nn_kernel<T, D, 0, true, MINB><<<blocks, kThreadsPerBlock, 0, c.st>>>(a);          \
    } else {                                                                               \
      if (rec)                                                                             \
        nn_kernel<T, D, 3, false, MINB><<<blocks, kThreadsPerBlock, 0, c.st>>>(a);         \
      else                                                                                 \
        nn_kernel<T, D, 0, false, MINB><<<blocks, kThreadsPerBlock, 0, c.st>>>(a);         \
    }                                                                                      \
  } while (0)
#define PICO_LAUNCH_NN(D)         \
  do {                            \
    if (mode & 16)                \
      PICO_LAUNCH_NN2(D, 12);     \
    else                          \
      PICO_LAUNCH_NN2(D, 16);     \
  } while (0)
      if (t->sdim == 2)
        PICO_LAUNCH_NN(2);
      else
        PICO_LAUNCH_NN(3);
#undef PICO_LAUNCH_NN
#undef PICO_LAUNCH_NN2
      PICO_CUDA(cudaGetLastError());
      if (use_fat) {
        KnnArgs<T> f = a;
        f.perm = a.tie_list;
        f.nq_from = a.tie_count;
        const unsigned fix_blocks = std::min<unsigned>(blocks, (unsigned)t->sm_count * 2);
        if (t->sdim == 2)
          knn_thread_kernel<T, 2, 1, true, false><<<fix_blocks, kThreadsPerBlock, 0, c.st>>>(f);
        else
          knn_thread_kernel<T, 3, 1, true, false><<<fix_blocks, kThreadsPerBlock, 0, c.st>>>(f);
        *launches += 1;
      }
    } else {
      switch (t->sdim) {
        case 1:
          launch_knn_thread<T, 1>(a, fast, deep, blocks, c.st);
          break;
        case 2:
          launch_knn_thread<T, 2>(a, fast, deep, blocks, c.st);
          break;
        default:
          launch_knn_thread<T, 3>(a, fast, deep, blocks, c.st);
          break;
      }
    }
  } else {
    unsigned blocks;
    size_t smem;
    PICO_TRY(warp_geometry<T>(c, t, nq, sizeof(WarpFrame<T>), plan_warp_smem<T>(t, a), &a.ws, &a.ws_depth, &blocks,
                              &smem));
    const bool reg = k <= 32;
#define PICO_LAUNCH_WARP(P, R)                                                                \
  do {                                                                                        \
    PICO_TRY(persistent_grid(c, t, knn_warp_kernel<T, P, R>, smem, &a.counter, &blocks));     \
    knn_warp_kernel<T, P, R><<<blocks, kWarpsPerBlock * 32, smem, c.st>>>(a);                 \
  } while (0)
    if (t->packed()) {
      if (reg)
        PICO_LAUNCH_WARP(true, true);
      else
        PICO_LAUNCH_WARP(true, false);
    } else {
      if (reg)
        PICO_LAUNCH_WARP(false, true);
      else
        PICO_LAUNCH_WARP(false, false);
    }
#undef PICO_LAUNCH_WARP
  }
  PICO_CUDA(cudaGetLastError());
  PICO_TRY(c.span_end());
  *launches += 1;
  PICO_TRY(c.mark(3));
  if (!on_device)
    PICO_CUDA(cudaMemcpyAsync(out, d_out, nq * k * sizeof(Neighbor<T>), cudaMemcpyDeviceToHost, c.st));
  PICO_TRY(c.mark(4));
  return 0;
}

constexpr size_t kHostChunk = (size_t)1 << 20;  // queries per pipelined chunk (host buffers)
constexpr int kHostStreams = 4;  // profiles/r2/host_pipeline_sweep_r2_v6.txt (round 1, before nn_tile_kernel: 6)
constexpr int kMaxHostStreams = 12;

// tuning hooks (profiles/r1/e2e_sweep.txt): PICO_B200_HOST_CHUNK, PICO_B200_HOST_STREAMS
size_t host_chunk() {
  static const size_t v = [] {
    const char* e = getenv("PICO_B200_HOST_CHUNK");
    const long long x = e ? atoll(e) : 0;
    return x >= 65536 ? (size_t)x : kHostChunk;
  }();
  return v;
}
int host_streams() {
  static const int v = [] {
    const char* e = getenv("PICO_B200_HOST_STREAMS");
    const int x = e ? atoi(e) : 0;
    return (x >= 1 && x <= kMaxHostStreams) ? x : kHostStreams;
  }();
  return v;
}
// The call ends one traversal + one D2H after the last chunk's H2D, so the tail of the batch is cut
// into halving chunks down to this many queries (0 = equal chunks). PICO_B200_HOST_TAPER is a tuning hook.
constexpr size_t kHostTaper = 262144;  // profiles/r2/host_pipeline_sweep_r2_v5.txt, _v6.txt: 2.23 -> 2.16 ms
size_t host_taper() {
  static const size_t v = [] {
    const char* e = getenv("PICO_B200_HOST_TAPER");
    const long long x = e ? atoll(e) : -1;
    return x >= 0 ? (size_t)x : kHostTaper;
  }();
  return v;
}
// Chunks whose H2D copy is issued (on a copy stream of its own) before the traversal of the current one is
// enqueued; 0 = every chunk copies on its own compute stream. PICO_B200_HOST_AHEAD is a tuning hook.
constexpr int kHostAhead = 64;  // all chunks (profiles/r1/host_pipeline_sweep.txt)
int host_ahead() {
  static const int v = [] {
    const char* e = getenv("PICO_B200_HOST_AHEAD");
    const int x = e ? atoi(e) : -1;
    return x >= 0 ? x : kHostAhead;
  }();
  return v;
}
// Order the chunks on a high-priority stream (PICO_B200_HOST_PRIO=0 keeps the ordering on the chunk's own stream).
bool host_priority_order() {
  static const bool v = [] {
    const char* e = getenv("PICO_B200_HOST_PRIO");
    return !(e && atoi(e) == 0);
  }();
  return v;
}
// PICO_B200_FUSED=1 (tuning hook): non-pipelined calls order and traverse in one kernel as well
bool resident_fused_order() {
  static const bool v = [] {
    const char* e = getenv("PICO_B200_FUSED");
    return e && atoi(e) != 0;
  }();
  return v;
}
// PICO_B200_HOST_FUSED=0 (tuning hook): chunks of the host pipeline never use nn_tile_kernel
bool host_fused_order() {
  static const bool v = [] {
    const char* e = getenv("PICO_B200_HOST_FUSED");
    return !(e && atoi(e) == 0);
  }();
  return v;
}
// PICO_B200_HOST_LOCAL_ORDER=1 (tuning hook): chunks of the host pipeline may use the tile-local order too
bool host_local_order() {
  static const bool v = [] {
    const char* e = getenv("PICO_B200_HOST_LOCAL_ORDER");
    return e && atoi(e) != 0;
  }();
  return v;
}
// Block slots per SM a chunk's traversal may occupy while the pipeline orders the next chunks (0 = all).
// PICO_B200_HOST_TRAV_BLOCKS is a tuning hook.
constexpr int kHostTraversalBlocks = 0;
int host_traversal_blocks() {
  static const int v = [] {
    const char* e = getenv("PICO_B200_HOST_TRAV_BLOCKS");
    const int x = e ? atoi(e) : -1;
    return (x >= 0 && x <= 16) ? x : kHostTraversalBlocks;
  }();
  return v;
}
constexpr size_t kHostHead = 131072;
size_t host_head() {
  static const size_t v = [] {
    const char* e = getenv("PICO_B200_HOST_HEAD");
    const long long x = e ? atoll(e) : -1;
    return x >= 0 ? (size_t)x : kHostHead;
  }();
  return v;
}
bool host_timeline() {
  static const bool v = getenv("PICO_B200_TIMELINE") != nullptr;  // per-chunk event times on stderr
  return v;
}

// [begin, count) of every chunk of a pipelined host batch
std::vector<std::pair<size_t, size_t>> host_chunk_plan(size_t nq, size_t chunk, size_t taper, size_t head) {
  std::vector<std::pair<size_t, size_t>> plan;
  size_t pos = 0;
  // a short first chunk lets the first traversal start early (the device idles during the first H2D)
  for (size_t c = head; c && c < chunk && nq - pos > 2 * chunk; c *= 2) {
    plan.emplace_back(pos, c);
    pos += c;
  }
  while (nq - pos > chunk && !(taper && nq - pos < 2 * chunk)) {
    plan.emplace_back(pos, chunk);
    pos += chunk;
  }
  if (taper) {
    while (nq - pos > 2 * taper) {
      const size_t c = ((nq - pos) / 2 + 4095) & ~(size_t)4095;
      plan.emplace_back(pos, c);
      pos += c;
    }
  }
  if (nq > pos) plan.emplace_back(pos, nq - pos);
  return plan;
}

}  // namespace

// pico_b200_free: a huge-page result block is kept for the next big result instead of going back to the OS
void release_result(void* p) {
  if (!p) return;
  void* drop = nullptr;
  {
    std::lock_guard<std::mutex> lock(g_results.mu);
    auto& big = g_results.big;
    auto it = std::find_if(big.begin(), big.end(), [p](const std::pair<void*, size_t>& b) { return b.first == p; });
    if (it == big.end()) {
      drop = p;
    } else if (it->second <= result_cache_budget() && it->second > g_results.spare_cap) {
      drop = g_results.spare;  // the smaller spare goes
      if (drop) big.erase(std::find_if(big.begin(), big.end(), [drop](const std::pair<void*, size_t>& b) { return b.first == drop; }));
      // (`it` may have moved: look the block up again)
      it = std::find_if(big.begin(), big.end(), [p](const std::pair<void*, size_t>& b) { return b.first == p; });
      g_results.spare = p;
      g_results.spare_cap = it->second;
    } else {
      big.erase(it);
      drop = p;
    }
  }
  free(drop);
}

// ------------------------------------------------------------------ knn
template <typename T>
int knn_batch(const pico_b200_tree* t, const T* q, size_t nq, size_t stride, size_t k, double e, Neighbor<T>* out,
              unsigned flags, pico_b200_search_stats* stats) {
  if (nq == 0 || k == 0) return 0;
  if (nq > 0x7fffffffu) return fail(PICO_B200_ERR_UNSUPPORTED, "more than 2^31-1 queries in one call (sort and scan counts are 32-bit)");
  if (k > 0x7fffffffu) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "k too large");
  const bool on_device = flags & PICO_B200_DEVICE_POINTERS;
  const bool want_async = (flags & PICO_B200_ASYNC) && on_device;
  uint64_t launches = 0;
  // Host buffers, big batch, no caller stream: split into chunks that rotate over a few
  // streams so that H2D of chunk i+1, traversal of chunk i and D2H of chunk i-1 overlap
  // (PCIe is full duplex). Each chunk is Z-ordered on its own.
  if (!on_device && !g_cfg.has_user_stream && nq >= 2 * host_chunk()) {
    const size_t chunk = host_chunk();
    const int n_streams = host_streams();
    const size_t sdim = t->sdim;
    // pageable buffers go through pinned mirrors (see PinnedMirror)
    bool stage_in = is_pageable(q), stage_out = is_pageable(out);
    if (stage_in || stage_out) {
      const size_t need = (stage_in ? nq * sdim * sizeof(T) : 0) + (stage_out ? nq * k * sizeof(Neighbor<T>) : 0);
      const bool have = (!stage_in || g_pin_in.cap >= nq * sdim * sizeof(T)) &&
                        (!stage_out || g_pin_out.cap >= nq * k * sizeof(Neighbor<T>));
      if (!have && (need > kMirrorBudget || ++g_pageable_batches < 2)) stage_in = stage_out = false;
    }
    // plain pageable output: each chunk's pages are faulted in (by several threads) right before its
    // copies are enqueued, while the device works on the previous chunks
    const bool touch_out = !stage_out && is_pageable(out);
    const T* src = q;
    size_t src_stride = stride;
    Neighbor<T>* dst = out;
    if (stage_in) {
      PICO_TRY(g_pin_in.reserve(nq * sdim * sizeof(T)));
      src = static_cast<const T*>(g_pin_in.p);
      src_stride = sdim;
    }
    if (stage_out) {
      PICO_TRY(g_pin_out.reserve(nq * k * sizeof(Neighbor<T>)));
      dst = static_cast<Neighbor<T>*>(g_pin_out.p);
    }
    CallCtx cp;  // copy stream + whole-batch device buffers (declared first: released after the compute streams)
    CallCtx ctx_all[kMaxHostStreams];
    CallCtx* ctx = ctx_all;
    for (int i = 0; i < n_streams; ++i) PICO_TRY(ctx[i].init(t->device));
    const auto plan = host_chunk_plan(nq, chunk, host_taper(), host_head());
    const size_t n_chunks = plan.size();
    // copy-ahead mode (pinned buffers): the queries of the next `ahead` chunks go up on one copy stream, each
    // followed by an event its compute stream waits for, so the H2D engine never idles behind the host thread
    // that is still enqueueing the ordering / traversal launches of earlier chunks. Pageable buffers keep the
    // per-chunk copies on the compute streams, interleaved with the host-side packing / page touching.
    // (whole-batch device buffers: only while they stay small against the HBM; huge batches keep per-chunk buffers)
    constexpr size_t kAheadBudget = (size_t)8 << 30;
    const size_t batch_bytes = nq * sdim * sizeof(T) + nq * k * sizeof(Neighbor<T>);
    const size_t ahead =
        (stage_in || stage_out || touch_out || batch_bytes > kAheadBudget) ? 0 : (size_t)host_ahead();
    T* d_q_all = nullptr;
    Neighbor<T>* d_out_all = nullptr;
    std::vector<cudaEvent_t> ready(ahead ? n_chunks : 0), ordered(ahead ? n_chunks : 0);
    // ordering of every chunk on the high-priority stream: one scratch area (the chunks are ordered one after
    // the other there) and one permutation array for the whole batch
    cudaStream_t hp = nullptr;
    PermPlan perm_plan;
    char* perm_scratch = nullptr;
    uint32_t* perm_all = nullptr;
    const int perm_bits = morton_bits(k == 1);
    // (the chunks always take the device-wide sort: the tile kernel's 96-register blocks wait longer for room
    // among the traversal blocks than the sort's kernels do — 2.68 against 2.46 ms, profiles/r2/order_sweep_v1.txt)
    // A batch the tree expects to be locally coherent (pico_b200_order_hint) is ordered INSIDE the traversal kernel,
    // tile by tile (nn_tile_kernel): no ordering kernels between a chunk's upload and its traversal. Every 16th call
    // (and while nothing is known yet) the first chunk is measured with tile_order_kernel on the side.
    const bool chunks_fused = host_fused_order() && k == 1 && order_locally(t);
    const bool chunks_local = !chunks_fused && host_local_order() && k == 1 && order_locally(t);
    const bool hp_order = ahead && host_priority_order() && !(flags & PICO_B200_NO_REORDER) && !chunks_local && !chunks_fused;
    const bool probe_first = k == 1 && order_mode() == 0 && !(flags & PICO_B200_NO_REORDER) && t->packed() &&
                             (t->order_hint.calls.fetch_add(1, std::memory_order_relaxed) % 16 == 0);
    if (ahead) {
      PICO_TRY(cp.init(t->device));
      cp.timed = false;
      PICO_TRY(cp.alloc(reinterpret_cast<void**>(&d_q_all), nq * sdim * sizeof(T)));
      PICO_TRY(cp.alloc(reinterpret_cast<void**>(&d_out_all), nq * k * sizeof(Neighbor<T>)));
      for (auto& ev : ready) PICO_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      for (auto& ev : ordered) PICO_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      if (hp_order) {
        size_t biggest = 0;
        for (auto& pc : plan) biggest = std::max(biggest, pc.second);
        PICO_TRY(priority_stream(t->device, &hp));
        PICO_TRY(plan_perm(biggest, perm_bits, &perm_plan));
        PICO_TRY(cp.alloc(reinterpret_cast<void**>(&perm_scratch), perm_plan.scratch_bytes()));
        PICO_TRY(cp.alloc(reinterpret_cast<void**>(&perm_all), nq * sizeof(uint32_t)));
      }
    }
    size_t uploaded = 0;  // chunks whose H2D has been issued
    auto upload_until = [&](size_t last) -> int {
      for (; uploaded <= last && uploaded < n_chunks; ++uploaded) {
        const size_t begin = plan[uploaded].first, cnt = plan[uploaded].second;
        if (src_stride == sdim)
          PICO_CUDA(cudaMemcpyAsync(d_q_all + begin * sdim, src + begin * sdim, cnt * sdim * sizeof(T),
                                    cudaMemcpyHostToDevice, cp.st));
        else
          PICO_CUDA(cudaMemcpy2DAsync(d_q_all + begin * sdim, sdim * sizeof(T), src + begin * src_stride,
                                      src_stride * sizeof(T), sdim * sizeof(T), cnt, cudaMemcpyHostToDevice, cp.st));
        PICO_CUDA(cudaEventRecord(ready[uploaded], cp.st));
      }
      return 0;
    };
    cudaEvent_t e0, e1;
    PICO_CUDA(cudaEventCreate(&e0));
    PICO_CUDA(cudaEventCreate(&e1));
    PICO_CUDA(cudaEventRecord(e0, ahead ? cp.st : ctx[0].st));
    const auto cpu0 = std::chrono::steady_clock::now();
    auto cpu_ms = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - cpu0).count(); };
    std::vector<double> cpu_at(2 * n_chunks, 0.0);
    std::vector<cudaEvent_t> done(stage_out ? n_chunks : 0);
    for (auto& ev : done) PICO_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    // results are copied out of the pinned mirror by a second thread while this one keeps
    // feeding chunks
    std::atomic<size_t> enqueued{0};
    std::atomic<bool> failed{false}, drain_failed{false};
    // Pageable buffers: ONE team of host threads that lives for the whole call. A single core of the host moves
    // ~7 GB/s, so the 86 MB of queries and 58 MB of results of the headline batch are 20 ms of one thread's time:
    // every worker PACKS query blocks into the pinned mirror while there are any (in chunk order, ahead of the chunk
    // loop, which waits for a chunk's blocks and lends a hand while it does) and otherwise DRAINS result blocks of
    // chunks whose device-to-host copy has completed, so the last chunk's results are moved by all of them.
    // (One team per chunk, spawned and joined around every chunk, kept the enqueueing thread busy 0.6-0.9 ms per
    // 1 Mi-query chunk: 6.5 ms per 7.2 M queries; a drain thread count sweep is in profiles/r2/pageable_pipeline.txt.)
    struct Block {
      uint32_t chunk;
      size_t begin, count;  // rows of the batch (packing) / bytes of the result (draining)
    };
    std::vector<Block> pack_blocks, drain_blocks;
    std::vector<std::atomic<uint32_t>> pack_left(stage_in ? n_chunks : 0);
    std::vector<std::atomic<uint8_t>> chunk_out(stage_out ? n_chunks : 0);  // 1: the chunk's results are in the mirror
    std::atomic<size_t> next_pack{0}, next_drain{0};
    if (stage_in) {
      const size_t rows_per_block = std::max<size_t>(1, ((size_t)768 << 10) / (sdim * sizeof(T)));
      for (size_t ci = 0; ci < n_chunks; ++ci) {
        uint32_t blocks = 0;
        for (size_t r = 0; r < plan[ci].second; r += rows_per_block, ++blocks)
          pack_blocks.push_back({(uint32_t)ci, plan[ci].first + r, std::min(rows_per_block, plan[ci].second - r)});
        pack_left[ci].store(blocks, std::memory_order_relaxed);
      }
    }
    if (stage_out) {
      const size_t bytes_per_block = (size_t)1 << 20;
      for (size_t ci = 0; ci < n_chunks; ++ci) {
        chunk_out[ci].store(0, std::memory_order_relaxed);
        const size_t first = plan[ci].first * k * sizeof(Neighbor<T>), bytes = plan[ci].second * k * sizeof(Neighbor<T>);
        for (size_t o = 0; o < bytes; o += bytes_per_block)
          drain_blocks.push_back({(uint32_t)ci, first + o, std::min(bytes_per_block, bytes - o)});
      }
    }
    auto pack_one = [&]() -> bool {
      if (next_pack.load(std::memory_order_relaxed) >= pack_blocks.size()) return false;
      const size_t b = next_pack.fetch_add(1, std::memory_order_relaxed);
      if (b >= pack_blocks.size()) return false;
      const Block& blk = pack_blocks[b];
      char* to = reinterpret_cast<char*>(const_cast<T*>(src) + blk.begin * sdim);
      const char* from = reinterpret_cast<const char*>(q + blk.begin * stride);
      if (stride == sdim)
        memcpy(to, from, blk.count * sdim * sizeof(T));
      else
        for (size_t r = 0; r < blk.count; ++r)
          memcpy(to + r * sdim * sizeof(T), from + r * stride * sizeof(T), sdim * sizeof(T));
      pack_left[blk.chunk].fetch_sub(1, std::memory_order_release);
      return true;
    };
    // has the chunk's device-to-host copy completed? (asked of the driver until it has, then remembered)
    auto chunk_ready = [&](uint32_t ci) -> bool {
      if (chunk_out[ci].load(std::memory_order_acquire)) return true;
      if (enqueued.load(std::memory_order_acquire) <= ci) return false;
      const cudaError_t st = cudaEventQuery(done[ci]);
      if (st == cudaSuccess) {
        chunk_out[ci].store(1, std::memory_order_release);
        return true;
      }
      if (st != cudaErrorNotReady) drain_failed.store(true);
      cudaGetLastError();
      return false;
    };
    // 1: moved a block; 0: the next block's chunk is not there yet; -1: nothing left to drain
    auto drain_one = [&]() -> int {
      const size_t peek = next_drain.load(std::memory_order_relaxed);
      if (peek >= drain_blocks.size()) return -1;
      if (!chunk_ready(drain_blocks[peek].chunk)) return 0;
      const size_t b = next_drain.fetch_add(1, std::memory_order_relaxed);
      if (b >= drain_blocks.size()) return -1;
      const Block& blk = drain_blocks[b];
      while (!chunk_ready(blk.chunk)) {  // (another worker took the block peeked at; this one belongs to a later chunk)
        if (failed.load() || drain_failed.load()) return -1;
        std::this_thread::yield();
      }
      memcpy(reinterpret_cast<char*>(out) + blk.begin, reinterpret_cast<const char*>(dst) + blk.begin, blk.count);
      return 1;
    };
    auto work = [&] {
      for (;;) {
        if (failed.load(std::memory_order_relaxed) || drain_failed.load(std::memory_order_relaxed)) return;
        if (stage_in && pack_one()) continue;
        if (!stage_out) return;
        const int d = drain_one();
        if (d < 0) return;
        if (d == 0) std::this_thread::yield();
      }
    };
    std::vector<std::thread> team;
    if (stage_in || stage_out) {
      static const unsigned workers = [] {
        const char* e = getenv("PICO_B200_HOST_WORKERS");  // tuning hook
        const int x = e ? atoi(e) : 0;
        if (x >= 1 && x <= 64) return (unsigned)x;
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        return std::max(1u, std::min(14u, hw > 2 ? hw - 2 : 1u));
      }();
      const int device = t->device;
      for (unsigned i = 0; i < workers; ++i)
        team.emplace_back([&, device] {
          cudaSetDevice(device);
          work();
        });
    }
    int rc = 0;
    const bool timeline = host_timeline() && n_chunks <= (size_t)n_streams;
    for (size_t ci = 0; ci < n_chunks && !rc; ++ci) {
      const size_t begin = plan[ci].first, cnt = plan[ci].second;
      CallCtx& c = ctx[ci % n_streams];
      c.timed = timeline;
      c.blocks_per_sm_cap = hp_order ? host_traversal_blocks() : 0;
      c.pipelined = !chunks_local;
      c.fused_order = chunks_fused;
      c.probe_order = probe_first && ci == 0;
      c.release();
      cpu_at[2 * ci] = cpu_ms();
      if (ahead) {
        rc = upload_until(ci + ahead - 1);
        const bool order_here = hp_order && cnt >= 2048;
        if (!rc && order_here) {
          if (cudaStreamWaitEvent(hp, ready[ci], 0) != cudaSuccess) rc = fail(PICO_B200_ERR_CUDA, "stream wait failed");
          if (!rc)
            rc = enqueue_perm<T>(hp, t, d_q_all + begin * sdim, sdim, cnt, perm_bits, perm_plan, perm_scratch,
                                 perm_all + begin);
          if (!rc && cudaEventRecord(ordered[ci], hp) != cudaSuccess) rc = fail(PICO_B200_ERR_CUDA, "event record failed");
        }
        if (!rc && cudaStreamWaitEvent(c.st, order_here ? ordered[ci] : ready[ci], 0) != cudaSuccess)
          rc = fail(PICO_B200_ERR_CUDA, "stream wait failed");

        if (!rc)
          rc = knn_enqueue<T>(c, t, d_q_all + begin * sdim, cnt, sdim, k, e, d_out_all + begin * k, flags, true,
                              &launches, order_here ? perm_all + begin : nullptr, order_here);
        if (!rc && cudaMemcpyAsync(dst + begin * k, d_out_all + begin * k, cnt * k * sizeof(Neighbor<T>),
                                   cudaMemcpyDeviceToHost, c.st) != cudaSuccess)
          rc = fail(PICO_B200_ERR_CUDA, "result copy failed");
        if (!rc) rc = c.mark(4);
      } else {
        if (stage_in)  // the chunk's queries must be in the mirror; pack blocks (of any chunk) while waiting
          while (pack_left[ci].load(std::memory_order_acquire) != 0)
            if (!pack_one()) std::this_thread::yield();
        if (touch_out) parallel_touch(reinterpret_cast<char*>(out + begin * k), cnt * k * sizeof(Neighbor<T>));
        rc = knn_enqueue<T>(c, t, src + begin * src_stride, cnt, src_stride, k, e, dst + begin * k, flags, false,
                            &launches);
      }
      if (!rc && stage_out && cudaEventRecord(done[ci], c.st) != cudaSuccess)
        rc = fail(PICO_B200_ERR_CUDA, "event record failed");
      if (rc) failed.store(true);
      enqueued.store(ci + 1, std::memory_order_release);
      cpu_at[2 * ci + 1] = cpu_ms();
    }
    if (!team.empty()) work();  // the enqueueing thread helps with what is left
    for (auto& th : team) th.join();
    for (auto& ev : done) cudaEventDestroy(ev);
    if (rc || drain_failed.load())  // quiesce before the events the streams wait on go away
      for (int i = 0; i < n_streams; ++i) cudaStreamSynchronize(ctx[i].st);
    if (rc || drain_failed.load()) {
      if (hp) cudaStreamSynchronize(hp);
      for (auto& ev : ready) cudaEventDestroy(ev);
      for (auto& ev : ordered) cudaEventDestroy(ev);
    }
    if (!rc && drain_failed.load()) rc = fail(PICO_B200_ERR_CUDA, "copying results out of the pinned mirror failed");
    if (rc) {
      cudaEventDestroy(e0);
      cudaEventDestroy(e1);
      return rc;
    }
    for (int i = 0; i < n_streams; ++i) PICO_CUDA(cudaStreamSynchronize(ctx[i].st));
    for (auto& ev : ready) cudaEventDestroy(ev);
    for (auto& ev : ordered) cudaEventDestroy(ev);
    PICO_CUDA(cudaEventRecord(e1, ctx[0].st));
    PICO_CUDA(cudaEventSynchronize(e1));
    if (timeline) {
      // device times of the chunk's stream (start = first work after its queries arrived when copying ahead),
      // and the host thread's clock when it began / finished enqueueing the chunk
      for (size_t i = 0; i < n_chunks; ++i)
        fprintf(stderr,
                "chunk %zu (%zu q): start %.3f h2d %.3f order %.3f traverse %.3f d2h %.3f ms | host enqueue %.3f-%.3f ms\n",
                i, plan[i].second, elapsed(e0, ctx[i].ev[0]), elapsed(e0, ctx[i].ev[1]), elapsed(e0, ctx[i].ev[2]),
                elapsed(e0, ctx[i].ev[3]), elapsed(e0, ctx[i].ev[4]), cpu_at[2 * i], cpu_at[2 * i + 1]);
      fprintf(stderr, "call: %.3f ms (copy-ahead %zu)\n", elapsed(e0, e1), ahead);
    }
    if (stats) {
      stats->h2d_ms = stats->reorder_ms = stats->d2h_ms = 0;
      stats->kernel_ms = elapsed(e0, e1);  // whole pipelined call
      stats->kernel_launches = launches;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return 0;
  }
  // (k <= 32: larger lists use the output row as working memory, which must not sit across PCIe)
  if (!on_device && !g_cfg.has_user_stream && !g_cfg.profiling && k <= 32 &&
      nq * t->sdim * sizeof(T) <= kSmallQueryBytes && nq * k * sizeof(Neighbor<T>) <= kSmallResultBytes) {
    PICO_TRY(g_small.ensure());
    T* mq = static_cast<T*>(g_small.p);
    Neighbor<T>* mout = reinterpret_cast<Neighbor<T>*>(static_cast<char*>(g_small.p) + kSmallQueryBytes);
    for (size_t i = 0; i < nq; ++i) memcpy(mq + i * t->sdim, q + i * stride, t->sdim * sizeof(T));
    CallCtx c;
    PICO_TRY(c.init(t->device));
    c.timed = stats != nullptr;
    PICO_TRY(knn_enqueue<T>(c, t, mq, nq, t->sdim, k, e, mout, flags | PICO_B200_NO_REORDER, true, &launches));
    PICO_CUDA(cudaStreamSynchronize(c.st));
    memcpy(out, mout, nq * k * sizeof(Neighbor<T>));
    if (stats) {
      stats->h2d_ms = stats->d2h_ms = stats->reorder_ms = 0;
      stats->kernel_ms = elapsed(c.ev[2], c.ev[3]);
      stats->kernel_launches = launches;
    }
    return 0;
  }
  CallCtx c;
  PICO_TRY(c.init(t->device, want_async));
  c.fused_order = resident_fused_order() && k == 1 && order_locally(t);
  PICO_TRY(knn_enqueue<T>(c, t, q, nq, stride, k, e, out, flags, on_device, &launches));
  if (c.async) return 0;
  PICO_CUDA(cudaStreamSynchronize(c.st));
  if (stats) {
    stats->h2d_ms = elapsed(c.ev[0], c.ev[1]);
    stats->reorder_ms = elapsed(c.ev[1], c.ev[2]);
    stats->kernel_ms = elapsed(c.ev[2], c.ev[3]);
    stats->d2h_ms = elapsed(c.ev[3], c.ev[4]);
    stats->kernel_launches = launches;
  }
  return 0;
}

// ------------------------------------------------------------------ radius
namespace {

// Sort every query's hits by distance: records are {int32 index, f32 distance}; read as
// little-endian uint64 the distance is the high word, and non-negative floats order like
// their bit patterns, so one segmented radix sort of 64-bit keys sorts by (distance, index).
int sort_hits_f32(CallCtx& c, Neighbor<float>* hits, size_t total, const uint64_t* d_offsets, size_t nq) {
  if (total == 0) return 0;
  if (total > 0x7fffffffu) return fail(PICO_B200_ERR_UNSUPPORTED, "too many hits to sort in one call");
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(hits);
  unsigned long long* keys2 = nullptr;
  PICO_TRY(c.alloc(reinterpret_cast<void**>(&keys2), total * 8));
  size_t tmp_bytes = 0;
  cub::DoubleBuffer<unsigned long long> db(keys, keys2);
  const unsigned long long* offs = reinterpret_cast<const unsigned long long*>(d_offsets);
  PICO_CUDA(cub::DeviceSegmentedSort::SortKeys(nullptr, tmp_bytes, db, (int)total, (int)nq, offs, offs + 1, c.st));
  void* tmp = nullptr;
  PICO_TRY(c.alloc(&tmp, tmp_bytes));
  PICO_CUDA(cub::DeviceSegmentedSort::SortKeys(tmp, tmp_bytes, db, (int)total, (int)nq, offs, offs + 1, c.st));
  if (db.Current() != keys)
    PICO_CUDA(cudaMemcpyAsync(keys, db.Current(), total * 8, cudaMemcpyDeviceToDevice, c.st));
  return 0;
}

// f64 records are 16 bytes {int32, pad, f64}: sort (distance bits, index) pairs.
__global__ void split_hits_f64(const Neighbor<double>* hits, size_t n, unsigned long long* keys, int* vals) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  keys[i] = (unsigned long long)__double_as_longlong(hits[i].distance);
  vals[i] = hits[i].index;
}
__global__ void join_hits_f64(Neighbor<double>* hits, size_t n, const unsigned long long* keys, const int* vals) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  hits[i].index = vals[i];
  hits[i].distance = __longlong_as_double((long long)keys[i]);
}

int sort_hits_f64(CallCtx& c, Neighbor<double>* hits, size_t total, const uint64_t* d_offsets, size_t nq) {
  if (total == 0) return 0;
  if (total > 0x7fffffffu) return fail(PICO_B200_ERR_UNSUPPORTED, "too many hits to sort in one call");
  unsigned long long *k1 = nullptr, *k2 = nullptr;
  int *v1 = nullptr, *v2 = nullptr;
  PICO_TRY(c.alloc(reinterpret_cast<void**>(&k1), total * 8));
  PICO_TRY(c.alloc(reinterpret_cast<void**>(&k2), total * 8));
  PICO_TRY(c.alloc(reinterpret_cast<void**>(&v1), total * 4));
  PICO_TRY(c.alloc(reinterpret_cast<void**>(&v2), total * 4));
  split_hits_f64<<<(unsigned)((total + 255) / 256), 256, 0, c.st>>>(hits, total, k1, v1);
  cub::DoubleBuffer<unsigned long long> dk(k1, k2);
  cub::DoubleBuffer<int> dv(v1, v2);
  const unsigned long long* offs = reinterpret_cast<const unsigned long long*>(d_offsets);
  size_t tmp_bytes = 0;
  PICO_CUDA(
      cub::DeviceSegmentedSort::SortPairs(nullptr, tmp_bytes, dk, dv, (int)total, (int)nq, offs, offs + 1, c.st));
  void* tmp = nullptr;
  PICO_TRY(c.alloc(&tmp, tmp_bytes));
  PICO_CUDA(cub::DeviceSegmentedSort::SortPairs(tmp, tmp_bytes, dk, dv, (int)total, (int)nq, offs, offs + 1, c.st));
  join_hits_f64<<<(unsigned)((total + 255) / 256), 256, 0, c.st>>>(hits, total, dk.Current(), dv.Current());
  PICO_CUDA(cudaGetLastError());
  return 0;
}

int sort_hits(CallCtx& c, Neighbor<float>* h, size_t total, const uint64_t* o, size_t nq) {
  return sort_hits_f32(c, h, total, o, nq);
}
int sort_hits(CallCtx& c, Neighbor<double>* h, size_t total, const uint64_t* o, size_t nq) {
  return sort_hits_f64(c, h, total, o, nq);
}

__global__ void widen_counts(const uint32_t* c, size_t n, uint64_t* out) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = c[i];
}

// counts[n] (u32) -> offsets[n+1] (u64, exclusive scan, last = total)
int scan_counts(CallCtx& c, const uint32_t* counts, size_t n, uint64_t** d_offsets) {
  uint64_t* wide = nullptr;
  uint64_t* offs = nullptr;
  PICO_TRY(c.alloc(reinterpret_cast<void**>(&wide), (n + 1) * 8));
  PICO_TRY(c.alloc(reinterpret_cast<void**>(&offs), (n + 1) * 8));
  PICO_CUDA(cudaMemsetAsync(wide, 0, (n + 1) * 8, c.st));
  widen_counts<<<(unsigned)((n + 255) / 256), 256, 0, c.st>>>(counts, n, wide);
  size_t tmp_bytes = 0;
  PICO_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, wide, offs, (int)(n + 1), c.st));
  void* tmp = nullptr;
  PICO_TRY(c.alloc(&tmp, tmp_bytes));
  PICO_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, wide, offs, (int)(n + 1), c.st));
  *d_offsets = offs;
  return 0;
}

template <typename T, bool FILL>
int launch_radius(CallCtx& c, const pico_b200_tree* t, RadiusArgs<T>& r, unsigned flags) {
  KnnArgs<T>& a = r.base;
  const bool use_thread = t->packed() && !(flags & PICO_B200_WARP_PER_QUERY);
  if (use_thread) {
    bool deep;
    unsigned blocks;
    PICO_TRY(thread_geometry(c, t, a, (int)t->sdim, &deep, &blocks));
#define PICO_RADIUS_T(D)                                                                          \
  do {                                                                                            \
    if (deep)                                                                                     \
      radius_thread_kernel<T, D, FILL, true><<<blocks, kThreadsPerBlock, 0, c.st>>>(r);           \
    else                                                                                          \
      radius_thread_kernel<T, D, FILL, false><<<blocks, kThreadsPerBlock, 0, c.st>>>(r);          \
  } while (0)
    switch (t->sdim) {
      case 1:
        PICO_RADIUS_T(1);
        break;
      case 2:
        PICO_RADIUS_T(2);
        break;
      default:
        PICO_RADIUS_T(3);
        break;
    }
#undef PICO_RADIUS_T
  } else {
    unsigned blocks;
    size_t smem;
    PICO_TRY(warp_geometry<T>(c, t, a.nq, sizeof(WarpFrame<T>), plan_warp_smem<T>(t, a), &a.ws, &a.ws_depth, &blocks,
                              &smem));
    if (t->packed()) {
      PICO_TRY(persistent_grid(c, t, radius_warp_kernel<T, true>, smem, &a.counter, &blocks));
      radius_warp_kernel<T, true><<<blocks, kWarpsPerBlock * 32, smem, c.st>>>(r);
    } else {
      PICO_TRY(persistent_grid(c, t, radius_warp_kernel<T, false>, smem, &a.counter, &blocks));
      radius_warp_kernel<T, false><<<blocks, kWarpsPerBlock * 32, smem, c.st>>>(r);
    }
  }
  PICO_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

// Small host batches (the reference's one-query-per-call pattern): queries, counts, offsets and hits
// live in the pinned device-mapped buffer; the prefix sum of at most 256 counts is host bookkeeping.
// *served tells whether the call was answered here; if not, the regular path takes it.
template <typename T>
int radius_small(const pico_b200_tree* t, const T* q, size_t nq, size_t stride, double radius, double e,
                 uint64_t* offsets_out, void** out, unsigned flags, bool* served) {
  *served = false;
  const size_t sdim = t->sdim;
  if (nq > 256 || nq * sdim * sizeof(T) > 8 * 1024 || (flags & PICO_B200_SORT_RESULTS) || g_cfg.has_user_stream ||
      g_cfg.profiling)
    return 0;
  PICO_TRY(g_small.ensure());
  char* base = static_cast<char*>(g_small.p);
  T* mq = reinterpret_cast<T*>(base);
  uint32_t* counts = reinterpret_cast<uint32_t*>(base + 8 * 1024);
  uint64_t* offs = reinterpret_cast<uint64_t*>(base + 10 * 1024);
  Neighbor<T>* hits = reinterpret_cast<Neighbor<T>*>(base + kSmallQueryBytes);
  for (size_t i = 0; i < nq; ++i) memcpy(mq + i * sdim, q + i * stride, sdim * sizeof(T));
  CallCtx c;
  PICO_TRY(c.init(t->device));
  c.timed = false;
  RadiusArgs<T> r;
  fill_base(r.base, t, mq, sdim, nq, nullptr, e);
  r.radius = e > 0 ? T(radius) * r.base.e_inv : T(radius);
  r.counts = counts;
  r.offsets = nullptr;
  r.hits = nullptr;
  PICO_TRY((launch_radius<T, false>(c, t, r, flags)));
  PICO_CUDA(cudaStreamSynchronize(c.st));
  uint64_t total = 0;
  for (size_t i = 0; i < nq; ++i) {
    offs[i] = total;
    total += counts[i];
  }
  offs[nq] = total;
  if (total * sizeof(Neighbor<T>) > kSmallResultBytes) return 0;
  Neighbor<T>* h = static_cast<Neighbor<T>*>(malloc((total ? total : 1) * sizeof(Neighbor<T>)));
  if (!h) return fail(PICO_B200_ERR_OUT_OF_MEMORY, "host allocation of radius results failed");
  if (total) {
    if (sizeof(T) == 8) memset(hits, 0, total * sizeof(Neighbor<T>));  // defined padding
    r.offsets = offs;
    r.hits = hits;
    r.base.ws = nullptr;
    int rc = launch_radius<T, true>(c, t, r, flags);
    if (!rc && cudaStreamSynchronize(c.st) != cudaSuccess) rc = fail(PICO_B200_ERR_CUDA, "small radius batch failed");
    if (rc) {
      free(h);
      return rc;
    }
    memcpy(h, hits, total * sizeof(Neighbor<T>));
  }
  memcpy(offsets_out, offs, (nq + 1) * sizeof(uint64_t));
  *out = h;
  *served = true;
  return 0;
}

template <typename T>
int radius_batch(const pico_b200_tree* t, const T* q, size_t nq, size_t stride, double radius, double e,
                 uint64_t* offsets_out, void** out, unsigned flags, pico_b200_search_stats* stats) {
  *out = nullptr;
  if (nq > 0x7fffffffu) return fail(PICO_B200_ERR_UNSUPPORTED, "more than 2^31-1 queries in one call (sort and scan counts are 32-bit)");
  // PICO_B200_DEVICE_POINTERS: queries and offsets_out are device pointers and *out receives a device
  // buffer (pico_b200_free_device); the call still synchronises once to learn the total.
  const bool on_device = flags & PICO_B200_DEVICE_POINTERS;
  if (!on_device && nq > 0 && !stats) {
    bool served = false;
    PICO_TRY(radius_small<T>(t, q, nq, stride, radius, e, offsets_out, out, flags, &served));
    if (served) return 0;
  }
  CallCtx c;
  PICO_TRY(c.init(t->device));
  if (nq == 0) {
    if (on_device)
      PICO_CUDA(cudaMemsetAsync(offsets_out, 0, 8, c.st));
    else
      offsets_out[0] = 0;
    return 0;
  }
  PICO_CUDA(cudaEventRecord(c.ev[0], c.st));
  const T* d_q = nullptr;
  size_t d_stride = 0;
  PICO_TRY(stage_queries(c, q, nq, stride, t->sdim, on_device, &d_q, &d_stride));
  PICO_CUDA(cudaEventRecord(c.ev[1], c.st));
  uint32_t* perm = nullptr;
  PICO_TRY(make_perm(c, t, d_q, d_stride, nq, flags, &perm));
  PICO_CUDA(cudaEventRecord(c.ev[2], c.st));

  RadiusArgs<T> r;
  fill_base(r.base, t, d_q, d_stride, nq, perm, e);
  // search_approximate_radius scales the radius once (search_visitor.hpp:261-267)
  r.radius = e > 0 ? T(radius) * r.base.e_inv : T(radius);
  r.offsets = nullptr;
  r.hits = nullptr;
  PICO_TRY(c.alloc(reinterpret_cast<void**>(&r.counts), nq * 4));
  PICO_TRY(c.span_begin());
  PICO_TRY((launch_radius<T, false>(c, t, r, flags)));
  PICO_TRY(c.span_end());
  uint64_t* d_offsets = nullptr;
  PICO_TRY(scan_counts(c, r.counts, nq, &d_offsets));
  uint64_t total64 = 0;
  if (on_device) {
    PICO_CUDA(cudaMemcpyAsync(offsets_out, d_offsets, (nq + 1) * 8, cudaMemcpyDeviceToDevice, c.st));
    PICO_CUDA(cudaMemcpyAsync(&total64, d_offsets + nq, 8, cudaMemcpyDeviceToHost, c.st));
  } else {
    PICO_CUDA(cudaMemcpyAsync(offsets_out, d_offsets, (nq + 1) * 8, cudaMemcpyDeviceToHost, c.st));
  }
  PICO_CUDA(cudaStreamSynchronize(c.st));
  const size_t total = on_device ? (size_t)total64 : (size_t)offsets_out[nq];
  const size_t bytes = (total ? total : 1) * sizeof(Neighbor<T>);
  Neighbor<T>* h_hits = nullptr;
  Neighbor<T>* d_hits = nullptr;
  if (on_device) {
    // from the stream-ordered pool (its release threshold keeps freed blocks): a 6 GB cudaMalloc + cudaFree per
    // call cost more than both traversal passes together (bench configs: 78 ms per call, 20 ms of kernels)
    PICO_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&d_hits), bytes, c.st));
  } else {
    h_hits = static_cast<Neighbor<T>*>(alloc_result(bytes));
    if (!h_hits) return fail(PICO_B200_ERR_OUT_OF_MEMORY, "host allocation of radius results failed");
  }
  auto give_up = [&](int rc) {
    free(h_hits);
    if (on_device) cudaFreeAsync(d_hits, c.st);
    return rc;
  };
  if (total) {
    if (!on_device) {
      int rc = c.alloc(reinterpret_cast<void**>(&d_hits), bytes);
      if (rc) return give_up(rc);
    }
    if (sizeof(T) == 8) cudaMemsetAsync(d_hits, 0, bytes, c.st);  // defined padding
    r.offsets = d_offsets;
    r.hits = d_hits;
    r.base.ws = nullptr;
    int rc = c.span_begin();
    if (!rc) rc = launch_radius<T, true>(c, t, r, flags);
    if (!rc) rc = c.span_end();
    if (!rc && (flags & PICO_B200_SORT_RESULTS)) rc = sort_hits(c, d_hits, total, d_offsets, nq);
    if (rc) return give_up(rc);
    PICO_CUDA(cudaEventRecord(c.ev[3], c.st));
    if (!on_device) {
      rc = copy_out(c.st, h_hits, d_hits, total * sizeof(Neighbor<T>));
      if (rc) return give_up(rc);
    }
  } else {
    PICO_CUDA(cudaEventRecord(c.ev[3], c.st));
  }
  PICO_CUDA(cudaEventRecord(c.ev[4], c.st));
  PICO_CUDA(cudaStreamSynchronize(c.st));
  *out = on_device ? static_cast<void*>(d_hits) : static_cast<void*>(h_hits);
  if (stats) {
    stats->h2d_ms = elapsed(c.ev[0], c.ev[1]);
    stats->reorder_ms = elapsed(c.ev[1], c.ev[2]);
    stats->kernel_ms = elapsed(c.ev[2], c.ev[3]);
    stats->d2h_ms = elapsed(c.ev[3], c.ev[4]);
    stats->kernel_launches = (perm ? 3 : 0) + 4 + ((flags & PICO_B200_SORT_RESULTS) ? 1 : 0);
  }
  return 0;
}

// ------------------------------------------------------------------ box
// Small host batches of boxes, like radius_small.
template <typename T>
int box_small(const pico_b200_tree* t, const T* mins, const T* maxs, size_t nb, size_t stride, uint64_t* offsets_out,
              int32_t** out, bool* served) {
  *served = false;
  const size_t sdim = t->sdim;
  if (nb > 256 || nb * sdim * sizeof(T) > 4 * 1024 || g_cfg.has_user_stream || g_cfg.profiling) return 0;
  PICO_TRY(g_small.ensure());
  char* base = static_cast<char*>(g_small.p);
  T* mmin = reinterpret_cast<T*>(base);
  T* mmax = reinterpret_cast<T*>(base + 4 * 1024);
  uint32_t* counts = reinterpret_cast<uint32_t*>(base + 8 * 1024);
  uint64_t* offs = reinterpret_cast<uint64_t*>(base + 10 * 1024);
  int32_t* hits = reinterpret_cast<int32_t*>(base + kSmallQueryBytes);
  for (size_t i = 0; i < nb; ++i) {
    memcpy(mmin + i * sdim, mins + i * stride, sdim * sizeof(T));
    memcpy(mmax + i * sdim, maxs + i * stride, sdim * sizeof(T));
  }
  CallCtx c;
  PICO_TRY(c.init(t->device));
  c.timed = false;
  BoxArgs<T> a;
  a.nodes = static_cast<const typename NodeOf<T>::type*>(t->d_nodes);
  a.outer = static_cast<const T*>(t->d_outer);
  a.metric = t->metric;
  a.pts4 = t->packed() ? static_cast<const typename Vec4Of<T>::type*>(t->d_pts) : nullptr;
  a.rows = t->packed() ? nullptr : static_cast<const T*>(t->d_pts);
  a.indices = t->d_indices;
  a.root_box = static_cast<const T*>(t->d_root_box);
  a.mins = mmin;
  a.maxs = mmax;
  a.stride = sdim;
  a.nb = (uint32_t)nb;
  a.sdim = (int)sdim;
  a.counts = counts;
  a.offsets = nullptr;
  a.hits = nullptr;
  unsigned blocks;
  size_t smem;
  PICO_TRY(warp_geometry<T>(c, t, nb, sizeof(BoxFrame) + sizeof(T), 4 * sdim * sizeof(T), &a.ws, &a.ws_depth, &blocks,
                            &smem));
  auto launch = [&]() -> int {
    if (t->packed()) {
      box_warp_kernel<T, true><<<blocks, kWarpsPerBlock * 32, smem, c.st>>>(a);
    } else {
      if (smem > 48 * 1024)
        PICO_CUDA(cudaFuncSetAttribute(box_warp_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
      box_warp_kernel<T, false><<<blocks, kWarpsPerBlock * 32, smem, c.st>>>(a);
    }
    PICO_CUDA(cudaGetLastError());
    PICO_CUDA(cudaStreamSynchronize(c.st));
    return 0;
  };
  PICO_TRY(launch());
  uint64_t total = 0;
  for (size_t i = 0; i < nb; ++i) {
    offs[i] = total;
    total += counts[i];
  }
  offs[nb] = total;
  if (total * sizeof(int32_t) > kSmallResultBytes) return 0;
  int32_t* h = static_cast<int32_t*>(malloc((total ? total : 1) * sizeof(int32_t)));
  if (!h) return fail(PICO_B200_ERR_OUT_OF_MEMORY, "host allocation of box results failed");
  if (total) {
    a.offsets = offs;
    a.hits = hits;
    const int rc = launch();
    if (rc) {
      free(h);
      return rc;
    }
    memcpy(h, hits, total * sizeof(int32_t));
  }
  memcpy(offsets_out, offs, (nb + 1) * sizeof(uint64_t));
  *out = h;
  *served = true;
  return 0;
}

template <typename T>
int box_batch(const pico_b200_tree* t, const T* mins, const T* maxs, size_t nb, size_t stride, uint64_t* offsets_out,
              int32_t** out, unsigned flags, pico_b200_search_stats* stats) {
  *out = nullptr;
  if (nb > 0x7fffffffu) return fail(PICO_B200_ERR_UNSUPPORTED, "more than 2^31-1 boxes in one call (scan counts are 32-bit)");
  const bool on_device = flags & PICO_B200_DEVICE_POINTERS;  // same convention as radius_batch
  if (!on_device && nb > 0 && !stats) {
    bool served = false;
    PICO_TRY(box_small<T>(t, mins, maxs, nb, stride, offsets_out, out, &served));
    if (served) return 0;
  }
  CallCtx c;
  PICO_TRY(c.init(t->device));
  if (nb == 0) {
    if (on_device)
      PICO_CUDA(cudaMemsetAsync(offsets_out, 0, 8, c.st));
    else
      offsets_out[0] = 0;
    return 0;
  }
  PICO_CUDA(cudaEventRecord(c.ev[0], c.st));
  const T *d_min = nullptr, *d_max = nullptr;
  size_t d_stride = 0;
  if (!on_device && maxs == mins + t->sdim && stride == 2 * t->sdim) {
    // (min, max) row pairs of one array — the layout of the reference's binding
    // (_pyco_tree/kd_tree.hpp:240-268): one contiguous copy instead of two strided ones
    T* buf = nullptr;
    PICO_TRY(c.alloc(reinterpret_cast<void**>(&buf), nb * stride * sizeof(T)));
    PICO_CUDA(cudaMemcpyAsync(buf, mins, nb * stride * sizeof(T), cudaMemcpyHostToDevice, c.st));
    d_min = buf;
    d_max = buf + t->sdim;
    d_stride = stride;
  } else {
    PICO_TRY(stage_queries(c, mins, nb, stride, t->sdim, on_device, &d_min, &d_stride));
    PICO_TRY(stage_queries(c, maxs, nb, stride, t->sdim, on_device, &d_max, &d_stride));
  }
  PICO_CUDA(cudaEventRecord(c.ev[1], c.st));
  PICO_CUDA(cudaEventRecord(c.ev[2], c.st));

  BoxArgs<T> a;
  a.nodes = static_cast<const typename NodeOf<T>::type*>(t->d_nodes);
  a.outer = static_cast<const T*>(t->d_outer);
  a.metric = t->metric;
  a.pts4 = t->packed() ? static_cast<const typename Vec4Of<T>::type*>(t->d_pts) : nullptr;
  a.rows = t->packed() ? nullptr : static_cast<const T*>(t->d_pts);
  a.indices = t->d_indices;
  a.root_box = static_cast<const T*>(t->d_root_box);
  a.mins = d_min;
  a.maxs = d_max;
  a.stride = d_stride;
  a.nb = (uint32_t)nb;
  a.sdim = (int)t->sdim;
  a.offsets = nullptr;
  a.hits = nullptr;
  PICO_TRY(c.alloc(reinterpret_cast<void**>(&a.counts), nb * 4));
  // thread-per-box for the packed euclidean case (the warp kernel: 98 ms per 1M boxes of 541 hits each — one box
  // per warp, frames in global memory, a __syncwarp per step; bench configs)
  const bool thread_box = t->packed() && !t->topological() && t->height < (size_t)kLocalStack - 1 &&
                          t->n_nodes < ((size_t)1 << 30) && !(flags & PICO_B200_WARP_PER_QUERY);
  uint32_t* box_perm = nullptr;
  if (thread_box) PICO_TRY(make_perm(c, t, d_min, d_stride, nb, flags, &box_perm));
  unsigned blocks = 0;
  size_t smem = 0;
  if (!thread_box)
    PICO_TRY(warp_geometry<T>(c, t, nb, sizeof(BoxFrame) + sizeof(T), 4 * t->sdim * sizeof(T), &a.ws, &a.ws_depth,
                              &blocks, &smem));
  auto launch = [&]() -> int {
    if (thread_box) {
      const unsigned tb = (unsigned)((nb + kThreadsPerBlock - 1) / kThreadsPerBlock);
      switch (t->sdim) {
        case 1:
          box_thread_kernel<T, 1><<<tb, kThreadsPerBlock, 0, c.st>>>(a, box_perm);
          break;
        case 2:
          box_thread_kernel<T, 2><<<tb, kThreadsPerBlock, 0, c.st>>>(a, box_perm);
          break;
        default:
          box_thread_kernel<T, 3><<<tb, kThreadsPerBlock, 0, c.st>>>(a, box_perm);
          break;
      }
      PICO_CUDA(cudaGetLastError());
      return 0;
    }
    if (t->packed()) {
      box_warp_kernel<T, true><<<blocks, kWarpsPerBlock * 32, smem, c.st>>>(a);
    } else {
      if (smem > 48 * 1024)
        PICO_CUDA(cudaFuncSetAttribute(box_warp_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
      box_warp_kernel<T, false><<<blocks, kWarpsPerBlock * 32, smem, c.st>>>(a);
    }
    PICO_CUDA(cudaGetLastError());
    return 0;
  };
  PICO_TRY(launch());
  uint64_t* d_offsets = nullptr;
  PICO_TRY(scan_counts(c, a.counts, nb, &d_offsets));
  uint64_t total64 = 0;
  if (on_device) {
    PICO_CUDA(cudaMemcpyAsync(offsets_out, d_offsets, (nb + 1) * 8, cudaMemcpyDeviceToDevice, c.st));
    PICO_CUDA(cudaMemcpyAsync(&total64, d_offsets + nb, 8, cudaMemcpyDeviceToHost, c.st));
  } else {
    PICO_CUDA(cudaMemcpyAsync(offsets_out, d_offsets, (nb + 1) * 8, cudaMemcpyDeviceToHost, c.st));
  }
  PICO_CUDA(cudaStreamSynchronize(c.st));
  const size_t total = on_device ? (size_t)total64 : (size_t)offsets_out[nb];
  const size_t bytes = (total ? total : 1) * sizeof(int32_t);
  int32_t* h_hits = nullptr;
  int32_t* d_hits = nullptr;
  if (on_device) {
    PICO_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&d_hits), bytes, c.st));
  } else {
    h_hits = static_cast<int32_t*>(alloc_result(bytes));
    if (!h_hits) return fail(PICO_B200_ERR_OUT_OF_MEMORY, "host allocation of box results failed");
  }
  auto give_up = [&](int rc) {
    free(h_hits);
    if (on_device) cudaFreeAsync(d_hits, c.st);
    return rc;
  };
  if (total) {
    int rc = on_device ? 0 : c.alloc(reinterpret_cast<void**>(&d_hits), bytes);
    if (!rc) {
      a.offsets = d_offsets;
      a.hits = d_hits;
      rc = launch();
    }
    if (rc) return give_up(rc);
    PICO_CUDA(cudaEventRecord(c.ev[3], c.st));
    if (!on_device) {
      rc = copy_out(c.st, h_hits, d_hits, total * sizeof(int32_t));
      if (rc) return give_up(rc);
    }
  } else {
    PICO_CUDA(cudaEventRecord(c.ev[3], c.st));
  }
  PICO_CUDA(cudaEventRecord(c.ev[4], c.st));
  PICO_CUDA(cudaStreamSynchronize(c.st));
  *out = on_device ? d_hits : h_hits;
  if (stats) {
    stats->h2d_ms = elapsed(c.ev[0], c.ev[1]);
    stats->reorder_ms = 0;
    stats->kernel_ms = elapsed(c.ev[2], c.ev[3]);
    stats->d2h_ms = elapsed(c.ev[3], c.ev[4]);
    stats->kernel_launches = 4;
  }
  return 0;
}

int order_state(const pico_b200_tree* t) {
  order_locally(t);
  return t->order_hint.state.load(std::memory_order_relaxed);
}

int set_thread_stream(void* stream, bool has) {
  g_cfg.user_stream = static_cast<cudaStream_t>(stream);
  g_cfg.has_user_stream = has;
  return 0;
}

int profile_begin() {
  for (auto& sp : g_cfg.spans) {
    cudaEventDestroy(sp.first);
    cudaEventDestroy(sp.second);
  }
  g_cfg.spans.clear();
  g_cfg.profiling = true;
  return 0;
}

int profile_end(double* ms, uint64_t* launches) {
  g_cfg.profiling = false;
  double total = 0;
  for (auto& sp : g_cfg.spans) {
    PICO_CUDA(cudaEventSynchronize(sp.second));
    float t = 0;
    PICO_CUDA(cudaEventElapsedTime(&t, sp.first, sp.second));
    total += t;
  }
  if (ms) *ms = total;
  if (launches) *launches = g_cfg.spans.size();
  for (auto& sp : g_cfg.spans) {
    cudaEventDestroy(sp.first);
    cudaEventDestroy(sp.second);
  }
  g_cfg.spans.clear();
  return 0;
}

// ------------------------------------------------------------------ leaf scan in isolation (measurement)
template <typename T>
int leaf_scan_profile(const pico_b200_tree* t, const T* d_q, size_t nq, size_t stride, Neighbor<T>* d_out, int repeats,
                      double* descend_ms, double* scan_ms, uint64_t* scan_bytes) {
  if (!t->packed() || t->metric != PICO_B200_METRIC_L2_SQUARED)
    return fail(PICO_B200_ERR_UNSUPPORTED, "leaf-scan profile: sdim <= 3 and metric_l2_squared only");
  if (nq == 0 || nq > 0x7fffffffu || repeats < 1) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "bad nq or repeats");
  CallCtx c;
  PICO_TRY(c.init(t->device));
  c.timed = false;
  uint32_t* perm = nullptr;
  PICO_TRY(make_perm(c, t, d_q, stride, nq, 0, &perm, true));
  KnnArgs<T> a;
  fill_base(a, t, d_q, stride, nq, perm, 0.0);
  a.out = d_out;
  a.k = 1;
  int2* ranges = nullptr;
  unsigned long long* n_streamed = nullptr;
  PICO_TRY(c.alloc(reinterpret_cast<void**>(&ranges), nq * sizeof(int2)));
  PICO_TRY(c.alloc(reinterpret_cast<void**>(&n_streamed), sizeof(unsigned long long)));
  const unsigned blocks = (unsigned)((nq + kThreadsPerBlock - 1) / kThreadsPerBlock);
  cudaEvent_t ev[3];
  for (auto& e : ev) PICO_CUDA(cudaEventCreate(&e));
  double ms_descend = 0, ms_scan = 0;
  unsigned long long streamed = 0;
  int rc = 0;
  for (int r = 0; r < repeats && !rc; ++r) {
    cudaMemsetAsync(n_streamed, 0, sizeof(unsigned long long), c.st);
    cudaEventRecord(ev[0], c.st);
    switch (t->sdim) {
      case 1:
        first_leaf_kernel<T, 1><<<blocks, kThreadsPerBlock, 0, c.st>>>(a, ranges, n_streamed);
        cudaEventRecord(ev[1], c.st);
        leaf_scan_kernel<T, 1><<<blocks, kThreadsPerBlock, 0, c.st>>>(a, ranges);
        break;
      case 2:
        first_leaf_kernel<T, 2><<<blocks, kThreadsPerBlock, 0, c.st>>>(a, ranges, n_streamed);
        cudaEventRecord(ev[1], c.st);
        leaf_scan_kernel<T, 2><<<blocks, kThreadsPerBlock, 0, c.st>>>(a, ranges);
        break;
      default:
        first_leaf_kernel<T, 3><<<blocks, kThreadsPerBlock, 0, c.st>>>(a, ranges, n_streamed);
        cudaEventRecord(ev[1], c.st);
        leaf_scan_kernel<T, 3><<<blocks, kThreadsPerBlock, 0, c.st>>>(a, ranges);
        break;
    }
    cudaEventRecord(ev[2], c.st);
    cudaMemcpyAsync(&streamed, n_streamed, sizeof(streamed), cudaMemcpyDeviceToHost, c.st);
    if (cudaStreamSynchronize(c.st) != cudaSuccess || cudaGetLastError() != cudaSuccess) {
      rc = fail(PICO_B200_ERR_CUDA, "leaf-scan profile kernels failed");
      break;
    }
    ms_descend += elapsed(ev[0], ev[1]);
    ms_scan += elapsed(ev[1], ev[2]);
  }
  for (auto& e : ev) cudaEventDestroy(e);
  if (rc) return rc;
  if (descend_ms) *descend_ms = ms_descend / repeats;
  if (scan_ms) *scan_ms = ms_scan / repeats;
  // per launch: slot -> query id (4 B, Z-ordered batches only), query (sdim scalars), range (8 B), the leaf's
  // vec4 records, one neighbour record
  if (scan_bytes)
    *scan_bytes = nq * ((perm ? 4 : 0) + t->sdim * sizeof(T) + sizeof(int2) + sizeof(Neighbor<T>)) +
                  streamed * 4 * sizeof(T);
  return 0;
}

template int leaf_scan_profile<float>(const pico_b200_tree*, const float*, size_t, size_t, Neighbor<float>*, int, double*,
                                      double*, uint64_t*);
template int leaf_scan_profile<double>(const pico_b200_tree*, const double*, size_t, size_t, Neighbor<double>*, int,
                                       double*, double*, uint64_t*);
template int knn_batch<float>(const pico_b200_tree*, const float*, size_t, size_t, size_t, double, Neighbor<float>*,
                              unsigned, pico_b200_search_stats*);
template int knn_batch<double>(const pico_b200_tree*, const double*, size_t, size_t, size_t, double,
                               Neighbor<double>*, unsigned, pico_b200_search_stats*);
template int radius_batch<float>(const pico_b200_tree*, const float*, size_t, size_t, double, double, uint64_t*,
                                 void**, unsigned, pico_b200_search_stats*);
template int radius_batch<double>(const pico_b200_tree*, const double*, size_t, size_t, double, double, uint64_t*,
                                  void**, unsigned, pico_b200_search_stats*);
template int box_batch<float>(const pico_b200_tree*, const float*, const float*, size_t, size_t, uint64_t*, int32_t**,
                              unsigned, pico_b200_search_stats*);
template int box_batch<double>(const pico_b200_tree*, const double*, const double*, size_t, size_t, uint64_t*,
                               int32_t**, unsigned, pico_b200_search_stats*);

}  // namespace pico
