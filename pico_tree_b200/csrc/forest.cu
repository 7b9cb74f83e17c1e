// forest.cu — kd_forest on the device: N randomized kd-trees over Householder-reflected copies of
// the point set, searched best-bin-first with a leaf budget per tree and ONE neighbour list shared by all trees.
//
// Replaces examples/pico_understory/pico_understory/kd_forest.hpp:15-138 (class kd_forest, search_nearest :91-120),
// internal/rkd_tree_builder.hpp:26-41 (one reflected copy + one ordinary build per tree),
// internal/rkd_tree_hh_data.hpp:51-90 (reflection y = x - 2 (r.x) r) and
// internal/kd_tree_priority_search.hpp:24-142 (priority_search_nearest_euclidean).
//
// Build: the points go to the device once; per tree a streaming kernel writes the reflected copy, and the ordinary
// device build (build.cu: sliding midpoint, max_leaf_size, bounds from the space) runs on it — with
// `keep_outer`, because the priority search reads all four bounds of kd_tree_node_topological. The copy itself is
// dropped; the tree keeps its points in leaf order like every other tree here.
//
// Search: one warp per query, persistent blocks. For every tree in order: reflect the query (the dot product is a
// serial chain in index order, one lane; the update is element-wise, all lanes), then the best-bin-first loop. A
// descent is warp-uniform (broadcast node loads); the far children met on the way are written to a per-warp path
// buffer, the leaf is scanned with the lane-per-point / staged row machinery of traverse_warp.cuh (same rounding as
// the sequential sum), and afterwards the recorded far children that pass `visitor.max() > distance`
// (kd_tree_priority_search.hpp:122) enter the queue — the reference's unwinding recursion tests them all against
// the same max() (see forest.cuh).
//
// The queue. At most `max_leaves_visited - leaves_visited` more pops can happen, in increasing order, so only that
// many smallest entries can ever be popped: the queue is a SORTED array bounded by the leaf budget (worst entry
// dropped when full) — shared memory for budgets up to kSharedQueue entries, global memory beyond. Insertion is
// warp-parallel (binary search + block shift); pop is O(1). Order: (distance, pre-order node id), the oracle's rule
// (the reference compares node addresses on equal distances, which no port can reproduce).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <random>
#include <type_traits>
#include <vector>

#include "forest.cuh"
#include "traverse_warp.cuh"

struct pico_b200_forest {
  int device = 0;
  int scalar = PICO_B200_F32;
  size_t n = 0, sdim = 0, max_leaf_size = 0, height = 0, max_nodes = 0, max_leaf_points = 0;
  std::vector<pico_b200_tree*> trees;
  void* d_rotations = nullptr;  // [n_trees][sdim]
  void* d_views = nullptr;      // ForestTreeDev[n_trees]
  std::vector<unsigned char> h_rotations;
  double build_ms = 0.0;
  int sm_count = pico::kSmBlocks;
};

namespace pico {
namespace {

constexpr int kForestWarps = 8;          // warps per block
constexpr int kSharedQueue = 512;        // queue entries per warp kept in shared memory

struct ForestTreeDev {
  const void* nodes;
  const void* outer;
  const void* pts;  // float4 / double4 records (sdim <= 3) or rows, both in leaf order
  const int32_t* indices;
};

template <typename T>
struct QEntry {
  T dist;
  uint32_t node;
};

// rkd_tree_hh_data::rotate_space (rkd_tree_hh_data.hpp:51-63): one thread per point, the dot product accumulated in
// index order exactly like rotate_point (:79-90).
template <typename T>
__global__ void reflect_points_kernel(const T* __restrict__ src, size_t n, int sdim, const T* __restrict__ r,
                                      T* __restrict__ dst) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const T* x = src + i * sdim;
  T* y = dst + i * sdim;
  T dot = T(0);
  for (int j = 0; j < sdim; ++j) dot = add_rn(dot, mul_rn(r[j], x[j]));
  dot = mul_rn(dot, T(2));
  for (int j = 0; j < sdim; ++j) y[j] = sub_rn(x[j], mul_rn(dot, r[j]));
}

// Sorted (descending) bounded queue of one warp; every member is called by all 32 lanes with the same arguments.
template <typename T>
struct WarpQueue {
  QEntry<T>* a;
  int n, cap;
  __device__ __forceinline__ static bool less(T d1, uint32_t n1, T d2, uint32_t n2) {
    return d1 < d2 || (!(d2 < d1) && n1 < n2);
  }
  __device__ __forceinline__ void clear() { n = 0; }
  __device__ __forceinline__ bool empty() const { return n == 0; }
  __device__ __forceinline__ QEntry<T> top() const { return a[n - 1]; }
  __device__ __forceinline__ void pop() { --n; }
  __device__ __forceinline__ void insert(T dist, uint32_t node) {
    const int lane = threadIdx.x & 31;
    // p = entries that stay in front of the new one (they are greater): first index whose entry is not greater
    int lo = 0, hi = n;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      const QEntry<T> e = a[mid];
      if (less(dist, node, e.dist, e.node))
        lo = mid + 1;
      else
        hi = mid;
    }
    const int p = lo;
    // every lane has finished reading before any lane writes (lanes of a warp may run apart: without this, lane 0
    // could store the new entry while a slower lane is still searching — racecheck found exactly that when p == 1)
    __syncwarp();
    if (n == cap) {
      if (p == 0) return;  // greater than everything kept: it can never be popped within the leaf budget
      for (int b = 1; b < p; b += 32) {  // drop a[0]: a[1 .. p) move one to the left
        const int i = b + lane;
        QEntry<T> v;
        if (i < p) v = a[i];
        __syncwarp();
        if (i < p) a[i - 1] = v;
        __syncwarp();
      }
      if (lane == 0) {
        a[p - 1].dist = dist;
        a[p - 1].node = node;
      }
      __syncwarp();
      return;
    }
    for (int h = n; h > p; h -= 32) {  // a[p .. n) move one to the right, from the end
      const int i = h - lane;
      QEntry<T> v;
      if (i > p) v = a[i - 1];
      __syncwarp();
      if (i > p) a[i] = v;
      __syncwarp();
    }
    if (lane == 0) {
      a[p].dist = dist;
      a[p].node = node;
    }
    __syncwarp();
    ++n;
  }
};

template <typename T>
struct ForestArgs {
  const ForestTreeDev* trees;
  int n_trees;
  const T* rotations;  // [n_trees][sdim]
  const T* q;
  size_t q_stride;
  uint32_t nq;
  Neighbor<T>* out;
  int k, sdim;
  unsigned long long max_leaves;
  int warp_smem;    // scalars of shared memory per warp
  int tile_rows;    // staged leaf tile (row storage), 0 = lane-per-point straight from global memory
  int queue_cap;    // entries
  QEntry<T>* queue_ws;  // global queues [warps][queue_cap], or nullptr: queue in shared memory
  QEntry<T>* path_ws;   // [warps][path_len]
  int path_len;
  unsigned long long* counter;
};

template <typename T, bool PACKED, bool REGLIST>
__global__ void __launch_bounds__(kForestWarps * 32) forest_knn_kernel(ForestArgs<T> a) {
  using NodeT = typename NodeOf<T>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  T* base = reinterpret_cast<T*>(smem_raw) + (size_t)w * a.warp_smem;
  T* q0 = base;               // the query as given
  T* sq = q0 + a.sdim;        // reflected into the current tree's space
  T* tile = sq + a.sdim;      // leaf tile (row storage)
  const int tile_scalars = a.tile_rows ? a.tile_rows * (a.sdim + (int)(16 / sizeof(T))) : 0;
  const size_t warp_global = (size_t)blockIdx.x * kForestWarps + w;
  WarpQueue<T> queue;
  queue.cap = a.queue_cap;
  queue.a = a.queue_ws ? a.queue_ws + warp_global * (size_t)a.queue_cap
                       : reinterpret_cast<QEntry<T>*>(tile + ((tile_scalars + 3) & ~3));
  QEntry<T>* path = a.path_ws + warp_global * (size_t)a.path_len;

  for (;;) {
    unsigned long long item = 0;
    if (lane == 0) item = atomicAdd(a.counter, 1ull);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= a.nq) return;
    const T* qp = a.q + (size_t)item * a.q_stride;
    for (int j = lane; j < a.sdim; j += 32) q0[j] = qp[j];
    __syncwarp();
    Neighbor<T>* row = a.out + (size_t)item * a.k;
    WarpVisitKnn<T, typename std::conditional<REGLIST, WarpKnnReg<T>, WarpKnnMem<T>>::type> vis;
    if constexpr (REGLIST)
      vis.list.init(a.k);
    else
      vis.list.init(row, a.k);

    for (int t = 0; t < a.n_trees; ++t) {
      const ForestTreeDev tv = a.trees[t];
      const NodeT* nodes = static_cast<const NodeT*>(tv.nodes);
      const T* outer = static_cast<const T*>(tv.outer);
      const T* r = a.rotations + (size_t)t * a.sdim;
      // kd_forest.hpp:103 -> rkd_tree_hh_data::rotate_point (:79-90)
      T dot = T(0);
      if (lane == 0) {
        for (int j = 0; j < a.sdim; ++j) dot = add_rn(dot, mul_rn(__ldg(r + j), q0[j]));
        dot = mul_rn(dot, T(2));
      }
      dot = __shfl_sync(0xffffffffu, dot, 0);
      for (int j = lane; j < a.sdim; j += 32) sq[j] = sub_rn(q0[j], mul_rn(dot, __ldg(r + j)));
      __syncwarp();

      PointSet<T, PACKED> ps;
      ps.pts4 = PACKED ? static_cast<const typename Vec4Of<T>::type*>(tv.pts) : nullptr;
      ps.rows = PACKED ? nullptr : static_cast<const T*>(tv.pts);
      ps.indices = tv.indices;
      ps.sdim = a.sdim;

      // priority_search_nearest_euclidean::operator(), kd_tree_priority_search.hpp:47-63
      unsigned long long leaves_visited = 0;
      queue.clear();
      queue.insert(T(0), 0u);
      while (!queue.empty()) {
        const QEntry<T> top = queue.top();
        if (leaves_visited >= a.max_leaves || vis.max() < top.dist) break;
        queue.pop();
        // ---- one descent (:66-125)
        uint32_t node = top.node;
        int n_path = 0;
        T na, nb;
        uint32_t right, sd;
        int lb, le;
        load_node(nodes, node, na, nb, right, sd, lb, le);
        while (sd != PICO_B200_LEAF) {
          T left_min, right_max;
          load_outer(outer, node, left_min, right_max);
          uint32_t first, second;
          T dist;
          forest::branch_step<T>(na, nb, left_min, right_max, sq[sd], top.dist, node, right, first, second, dist);
          if (lane == 0) {
            path[n_path].node = second;
            path[n_path].dist = dist;
          }
          ++n_path;
          node = first;
          load_node(nodes, node, na, nb, right, sd, lb, le);
        }
        // ---- leaf (:68-73), candidates offered in leaf order
        if (!PACKED && a.tile_rows > 0) {
          scan_leaf_staged<T>(ps.rows, ps.indices, a.sdim, lb, le, sq, tile, a.tile_rows,
                              (int)PICO_B200_METRIC_L2_SQUARED, false, T(1), vis);
        } else {
          for (int b0 = lb; b0 < le; b0 += 32) {
            const int i = b0 + lane;
            int idx = -1;
            T d = Limits<T>::max();
            if (i < le) d = ps.distance(i, sq, (int)PICO_B200_METRIC_L2_SQUARED, idx);
            vis.visit_batch(i < le, idx, d);
          }
        }
        // ---- the far children of this descent that are still closer than max() (:122-124)
        __syncwarp();
        const T reach = vis.max();
        for (int b0 = 0; b0 < n_path; b0 += 32) {
          const int i = b0 + lane;
          QEntry<T> e;
          e.dist = T(0);
          e.node = 0;
          if (i < n_path) e = path[i];
          unsigned pass = __ballot_sync(0xffffffffu, i < n_path && reach > e.dist);
          while (pass) {
            const int l = __ffs(pass) - 1;
            pass &= pass - 1;
            queue.insert(__shfl_sync(0xffffffffu, e.dist, l), __shfl_sync(0xffffffffu, e.node, l));
          }
        }
        __syncwarp();
        ++leaves_visited;
      }
    }
    if constexpr (REGLIST) vis.list.store(row);
    __syncwarp();
  }
}

int forest_fail(int code, const std::string& msg) { return fail(code, msg); }

template <typename T>
int build_forest(pico_b200_forest* f, const T* h_pts, size_t stride, const T* rotations, size_t n_trees) {
  const size_t n = f->n, sdim = f->sdim;
  PICO_CUDA(cudaSetDevice(f->device));
  cudaDeviceProp prop;
  PICO_CUDA(cudaGetDeviceProperties(&prop, f->device));
  f->sm_count = prop.multiProcessorCount;
  // reflection vectors: given (tests, reproducible forests) or drawn like rkd_tree_hh_data::random_rotation
  // (rkd_tree_hh_data.hpp:14-29: i.i.d. N(0,1) coordinates, normalised)
  std::vector<T> rot(n_trees * sdim);
  if (rotations) {
    std::copy(rotations, rotations + n_trees * sdim, rot.begin());
  } else {
    std::random_device rd;
    for (size_t t = 0; t < n_trees; ++t) {
      std::mt19937 e(rd());
      std::normal_distribution<T> gaussian(T(0), T(1));
      T* v = rot.data() + t * sdim;
      T s = T(0);
      for (size_t i = 0; i < sdim; ++i) {
        v[i] = gaussian(e);
        s += v[i] * v[i];
      }
      s = std::sqrt(s);
      if (s > T(0))
        for (size_t i = 0; i < sdim; ++i) v[i] /= s;
      else
        v[0] = T(1);
    }
  }
  f->h_rotations.assign(reinterpret_cast<unsigned char*>(rot.data()),
                        reinterpret_cast<unsigned char*>(rot.data()) + rot.size() * sizeof(T));
  PICO_CUDA(cudaMalloc(&f->d_rotations, rot.size() * sizeof(T)));
  PICO_CUDA(cudaMemcpy(f->d_rotations, rot.data(), rot.size() * sizeof(T), cudaMemcpyHostToDevice));

  T *d_src = nullptr, *d_rot_pts = nullptr;
  PICO_CUDA(cudaMalloc(reinterpret_cast<void**>(&d_src), n * sdim * sizeof(T)));
  struct Free {
    T*& p;
    ~Free() { cudaFree(p); }
  } free_src{d_src};
  PICO_CUDA(cudaMalloc(reinterpret_cast<void**>(&d_rot_pts), n * sdim * sizeof(T)));
  Free free_rot{d_rot_pts};
  if (stride == sdim)
    PICO_CUDA(cudaMemcpy(d_src, h_pts, n * sdim * sizeof(T), cudaMemcpyDefault));
  else
    PICO_CUDA(cudaMemcpy2D(d_src, sdim * sizeof(T), h_pts, stride * sizeof(T), sdim * sizeof(T), n, cudaMemcpyDefault));

  cudaEvent_t e0, e1;
  PICO_CUDA(cudaEventCreate(&e0));
  PICO_CUDA(cudaEventCreate(&e1));
  PICO_CUDA(cudaEventRecord(e0, nullptr));
  std::vector<ForestTreeDev> views(n_trees);
  for (size_t t = 0; t < n_trees; ++t) {
    reflect_points_kernel<T><<<(unsigned)((n + 127) / 128), 128>>>(d_src, n, (int)sdim,
                                                                  static_cast<const T*>(f->d_rotations) + t * sdim,
                                                                  d_rot_pts);
    PICO_CUDA(cudaGetLastError());
    PICO_CUDA(cudaDeviceSynchronize());
    pico_b200_tree* tree = new (std::nothrow) pico_b200_tree();
    if (!tree) return forest_fail(PICO_B200_ERR_OUT_OF_MEMORY, "host allocation failed");
    f->trees.push_back(tree);
    tree->device = f->device;
    tree->scalar = f->scalar;
    tree->metric = PICO_B200_METRIC_L2_SQUARED;
    tree->n = n;
    tree->sdim = sdim;
    tree->sm_count = f->sm_count;
    tree->keep_outer = true;
    // kd_forest's constructor: max_leaf_size_t, bounds_from_space, sliding_midpoint_max_side (kd_forest.hpp:44-52)
    PICO_TRY(build_tree<T>(tree, d_rot_pts, sdim, PICO_B200_RULE_SLIDING_MIDPOINT_MAX_SIDE, PICO_B200_STOP_MAX_LEAF_SIZE,
                           f->max_leaf_size, nullptr, nullptr));
    f->height = std::max(f->height, tree->height);
    f->max_nodes = std::max(f->max_nodes, tree->n_nodes);
    f->max_leaf_points = std::max(f->max_leaf_points, tree->max_leaf_points);
    views[t].nodes = tree->d_nodes;
    views[t].outer = tree->d_outer;
    views[t].pts = tree->d_pts;
    views[t].indices = tree->d_indices;
  }
  PICO_CUDA(cudaEventRecord(e1, nullptr));
  PICO_CUDA(cudaEventSynchronize(e1));
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  f->build_ms = ms;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  PICO_CUDA(cudaMalloc(&f->d_views, n_trees * sizeof(ForestTreeDev)));
  PICO_CUDA(cudaMemcpy(f->d_views, views.data(), n_trees * sizeof(ForestTreeDev), cudaMemcpyHostToDevice));
  return 0;
}

template <typename T>
int forest_knn(const pico_b200_forest* f, const T* q, size_t nq, size_t stride, size_t k, size_t max_leaves,
               Neighbor<T>* out, unsigned flags, pico_b200_search_stats* stats) {
  if (nq == 0 || k == 0) return 0;
  if (nq > 0x7fffffffu) return forest_fail(PICO_B200_ERR_UNSUPPORTED, "more than 2^31-1 queries in one call");
  if (k > 0x7fffffffu) return forest_fail(PICO_B200_ERR_INVALID_ARGUMENT, "k too large");
  PICO_CUDA(cudaSetDevice(f->device));
  const bool on_device = flags & PICO_B200_DEVICE_POINTERS;
  const size_t sdim = f->sdim;
  cudaStream_t st;
  PICO_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  struct Guard {
    cudaStream_t s;
    std::vector<void*> bufs;
    ~Guard() {
      cudaStreamSynchronize(s);
      for (void* p : bufs) cudaFree(p);
      cudaStreamDestroy(s);
    }
  } g{st, {}};
  auto dalloc = [&](void** p, size_t bytes) -> int {
    PICO_CUDA(cudaMalloc(p, bytes ? bytes : 16));
    g.bufs.push_back(*p);
    return 0;
  };
  cudaEvent_t ev[4];
  for (auto& e : ev) PICO_CUDA(cudaEventCreate(&e));
  struct EvGuard {
    cudaEvent_t* e;
    ~EvGuard() {
      for (int i = 0; i < 4; ++i) cudaEventDestroy(e[i]);
    }
  } eg{ev};
  PICO_CUDA(cudaEventRecord(ev[0], st));
  const T* d_q = q;
  size_t d_stride = stride;
  Neighbor<T>* d_out = out;
  if (!on_device) {
    T* buf = nullptr;
    PICO_TRY(dalloc(reinterpret_cast<void**>(&buf), nq * sdim * sizeof(T)));
    if (stride == sdim)
      PICO_CUDA(cudaMemcpyAsync(buf, q, nq * sdim * sizeof(T), cudaMemcpyHostToDevice, st));
    else
      PICO_CUDA(cudaMemcpy2DAsync(buf, sdim * sizeof(T), q, stride * sizeof(T), sdim * sizeof(T), nq,
                                  cudaMemcpyHostToDevice, st));
    d_q = buf;
    d_stride = sdim;
    PICO_TRY(dalloc(reinterpret_cast<void**>(&d_out), nq * k * sizeof(Neighbor<T>)));
  }
  PICO_CUDA(cudaEventRecord(ev[1], st));

  ForestArgs<T> a;
  a.trees = static_cast<const ForestTreeDev*>(f->d_views);
  a.n_trees = (int)f->trees.size();
  a.rotations = static_cast<const T*>(f->d_rotations);
  a.q = d_q;
  a.q_stride = d_stride;
  a.nq = (uint32_t)nq;
  a.out = d_out;
  a.k = (int)k;
  a.sdim = (int)sdim;
  a.max_leaves = max_leaves;
  // a queue entry can only be popped while the leaf budget lasts, and a tree has max_nodes nodes to offer
  const size_t cap = std::max<size_t>(1, std::min<size_t>(max_leaves, f->max_nodes));
  a.queue_cap = (int)std::min<size_t>(cap, 0x7fffffff);
  const bool packed = sdim <= (size_t)kMaxPackedDim;
  const size_t vec = 16 / sizeof(T);
  a.tile_rows = 0;
  if (!packed && sdim % vec == 0 && f->max_leaf_points > 0) {
    size_t rows = std::min<size_t>(f->max_leaf_points, 32);
    while (rows > 1 && rows * (sdim + vec) * sizeof(T) > 12 * 1024) --rows;
    if (rows * (sdim + vec) * sizeof(T) <= 12 * 1024) a.tile_rows = (int)rows;
  }
  const bool shared_queue = cap <= (size_t)kSharedQueue;
  const size_t tile_scalars = a.tile_rows ? (size_t)a.tile_rows * (sdim + vec) : 0;
  size_t warp_scalars = 2 * sdim + ((tile_scalars + 3) & ~(size_t)3);
  warp_scalars = (warp_scalars + 3) & ~(size_t)3;
  // (the queue starts 16-byte aligned: warp_scalars and the offsets in front of it are multiples of four scalars)
  const size_t queue_scalars = shared_queue ? (cap * sizeof(QEntry<T>) + sizeof(T) - 1) / sizeof(T) : 0;
  warp_scalars += (queue_scalars + 3) & ~(size_t)3;
  // keep 2*sdim a multiple of four scalars so that everything behind it stays aligned
  if ((2 * sdim) % 4) warp_scalars += 4;
  a.warp_smem = (int)warp_scalars;
  const size_t smem = warp_scalars * sizeof(T) * kForestWarps;
  if (smem > 200 * 1024) return forest_fail(PICO_B200_ERR_UNSUPPORTED, "spatial dimension too large for shared memory");
  const bool reg = k <= 32;
  auto kernel = packed ? (reg ? forest_knn_kernel<T, true, true> : forest_knn_kernel<T, true, false>)
                       : (reg ? forest_knn_kernel<T, false, true> : forest_knn_kernel<T, false, false>);
  if (smem > 48 * 1024) PICO_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  PICO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kForestWarps * 32, smem));
  if (per_sm < 1) return forest_fail(PICO_B200_ERR_UNSUPPORTED, "forest kernel does not fit on an SM");
  size_t blocks = std::min<size_t>((nq + kForestWarps - 1) / kForestWarps, (size_t)per_sm * f->sm_count);
  a.path_len = (int)f->height + 2;
  a.queue_ws = nullptr;
  if (!shared_queue) {
    // global queues: bound the workspace (and with it the number of resident warps) to 4 GiB
    const size_t per_warp = cap * sizeof(QEntry<T>);
    const size_t budget = (size_t)4 << 30;
    blocks = std::max<size_t>(1, std::min(blocks, budget / (per_warp * kForestWarps)));
    if (per_warp * kForestWarps > budget)
      return forest_fail(PICO_B200_ERR_OUT_OF_MEMORY, "leaf budget too large for the queue workspace");
    PICO_TRY(dalloc(reinterpret_cast<void**>(&a.queue_ws), blocks * kForestWarps * per_warp));
  }
  PICO_TRY(dalloc(reinterpret_cast<void**>(&a.path_ws), blocks * kForestWarps * (size_t)a.path_len * sizeof(QEntry<T>)));
  PICO_TRY(dalloc(reinterpret_cast<void**>(&a.counter), sizeof(unsigned long long)));
  PICO_CUDA(cudaMemsetAsync(a.counter, 0, sizeof(unsigned long long), st));
  // the k > 32 list works in the output row: fill happens in the kernel (WarpKnnMem::init)
  kernel<<<(unsigned)blocks, kForestWarps * 32, smem, st>>>(a);
  PICO_CUDA(cudaGetLastError());
  PICO_CUDA(cudaEventRecord(ev[2], st));
  if (!on_device) PICO_CUDA(cudaMemcpyAsync(out, d_out, nq * k * sizeof(Neighbor<T>), cudaMemcpyDeviceToHost, st));
  PICO_CUDA(cudaEventRecord(ev[3], st));
  PICO_CUDA(cudaStreamSynchronize(st));
  if (stats) {
    float x = 0;
    cudaEventElapsedTime(&x, ev[0], ev[1]);
    stats->h2d_ms = x;
    stats->reorder_ms = 0;
    cudaEventElapsedTime(&x, ev[1], ev[2]);
    stats->kernel_ms = x;
    cudaEventElapsedTime(&x, ev[2], ev[3]);
    stats->d2h_ms = x;
    stats->kernel_launches = 1;
  }
  return 0;
}

}  // namespace
}  // namespace pico

using namespace pico;

extern "C" {

void pico_b200_forest_destroy(pico_b200_forest* f) {
  if (!f) return;
  cudaSetDevice(f->device);
  for (pico_b200_tree* t : f->trees) pico_b200_tree_destroy(t);
  cudaFree(f->d_rotations);
  cudaFree(f->d_views);
  delete f;
}

int pico_b200_forest_create(const void* pts, size_t n, size_t sdim, size_t stride, int scalar, size_t max_leaf_size,
                            const void* rotations, size_t forest_size, int device, pico_b200_forest** out) {
  if (!out) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "out is null");
  *out = nullptr;
  if (!pts || n == 0) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "empty point set");
  if (sdim == 0 || stride < sdim) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "bad sdim / stride");
  if (sdim > 0x7fff) return fail(PICO_B200_ERR_UNSUPPORTED, "sdim > 32767");
  if (scalar != PICO_B200_F32 && scalar != PICO_B200_F64) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "unknown scalar");
  if (max_leaf_size == 0) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "max_leaf_size must be > 0");
  if (forest_size == 0) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "forest_size must be > 0");
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
    cudaGetLastError();
    return fail(PICO_B200_ERR_NO_DEVICE, "no such CUDA device");
  }
  pico_b200_forest* f = new (std::nothrow) pico_b200_forest();
  if (!f) return fail(PICO_B200_ERR_OUT_OF_MEMORY, "host allocation failed");
  f->device = device;
  f->scalar = scalar;
  f->n = n;
  f->sdim = sdim;
  f->max_leaf_size = max_leaf_size;
  const int rc = scalar == PICO_B200_F32
                     ? build_forest<float>(f, static_cast<const float*>(pts), stride,
                                           static_cast<const float*>(rotations), forest_size)
                     : build_forest<double>(f, static_cast<const double*>(pts), stride,
                                            static_cast<const double*>(rotations), forest_size);
  if (rc) {
    pico_b200_forest_destroy(f);
    return rc;
  }
  *out = f;
  return 0;
}

int pico_b200_forest_info_get(const pico_b200_forest* f, pico_b200_forest_info* info) {
  if (!f || !info) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "null argument");
  info->n_points = f->n;
  info->sdim = f->sdim;
  info->n_trees = f->trees.size();
  info->max_leaf_size = f->max_leaf_size;
  info->height = f->height;
  info->scalar = f->scalar;
  info->device = f->device;
  info->build_ms = f->build_ms;
  uint64_t bytes = 0;
  for (const pico_b200_tree* t : f->trees) bytes += t->device_bytes;
  info->device_bytes = bytes;
  return 0;
}

int pico_b200_forest_rotations(const pico_b200_forest* f, void* out) {
  if (!f || !out) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "null argument");
  memcpy(out, f->h_rotations.data(), f->h_rotations.size());
  return 0;
}

int pico_b200_forest_tree(const pico_b200_forest* f, size_t i, const pico_b200_tree** tree) {
  if (!f || !tree) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "null argument");
  if (i >= f->trees.size()) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "tree index out of range");
  *tree = f->trees[i];
  return 0;
}

int pico_b200_forest_knn(const pico_b200_forest* f, const void* queries, size_t nq, size_t stride, size_t k,
                         size_t max_leaves_visited, void* neighbors_out, unsigned flags, pico_b200_search_stats* stats) {
  if (!f) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "forest is null");
  if (stats) memset(stats, 0, sizeof(*stats));
  if (nq && (!queries || !neighbors_out)) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "null argument");
  if (stride < f->sdim) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "stride smaller than sdim");
  if (k > f->n) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "k larger than the number of points");
  if (f->scalar == PICO_B200_F32)
    return forest_knn<float>(f, static_cast<const float*>(queries), nq, stride, k, max_leaves_visited,
                             static_cast<Neighbor<float>*>(neighbors_out), flags, stats);
  return forest_knn<double>(f, static_cast<const double*>(queries), nq, stride, k, max_leaves_visited,
                            static_cast<Neighbor<double>*>(neighbors_out), flags, stats);
}

}  // extern "C"
