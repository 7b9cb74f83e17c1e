// build.cu — device-side KdTree construction (sm_100a).
//
// Replaces the reference's recursive, single-threaded builder
// (internal/kd_tree_builder.hpp:351-396 create_node, :143-280 splitters, :471-494 driver)
// by a level-synchronous top-down pass over a BFS node table followed by a bottom-up pass
// for the tight child bounds and a pre-order renumbering:
//
//   1. root box      : min/max reduction over all points   (space_wrapper.hpp:34-40)
//   2. per level     : every node of the level is handled by one cooperating group —
//                      a warp (count <= kWarpNodeMax) or a whole CTA (big nodes, listed
//                      separately) — which either finishes it as a leaf (tight box,
//                      kd_tree_builder.hpp:399-407) or splits it: max_side of the CUT box
//                      (box.hpp:71-82), split_val, partition of the index range, slide
//                      fix-up, two children appended to the next level.
//   3. bottom-up     : tight box of a branch = fit(left, right) (:392); left_max/right_min
//                      (kd_tree_node.hpp:86-92); subtree sizes.
//   4. top-down      : pre-order numbers; emit 16-B nodes; gather points into leaf order.
//
// The index array is partitioned in place exactly like libstdc++'s std::partition would
// leave it (the j-th misplaced element from the left is exchanged with the j-th misplaced
// element from the right), so the order inside a leaf — which decides which of two
// equidistant points a query reports — matches the reference. Slides and the median rule call
// std::nth_element in the reference; group_nth_element emulates libstdc++'s introselect exactly,
// so the index permutation (and with it the saved stream) is identical to the reference's.
#include <cub/cub.cuh>

#include <algorithm>
#include <chrono>
#include <vector>

#include "common.cuh"

namespace pico {
namespace {

constexpr int kWarpNodeMax = 1024;  // nodes up to this many points are split by one warp
constexpr int kBigThreads = 1024;   // CTA size for bigger nodes
// Nodes above kHugeMin points (the first ~8 levels of a multi-million point tree) are not given
// to a single CTA: all such nodes of a level are cut into chunks of kChunk index positions and
// split by grid-wide passes (plan, count, resolve, [slide], scatter, swap) so that every SM works
// on them. profiles/r1/launches_knn1_v2.csv: one CTA per node spent 38 of 45 ms there.
constexpr int kHugeMin = 32768;
constexpr int kChunkThreads = 256;
constexpr int kChunkItems = 8;
constexpr int kChunk = kChunkThreads * kChunkItems;

template <typename T>
struct BNode {
  int32_t begin, end;
  int32_t left, right;   // BFS ids, -1 for a leaf
  int32_t split_dim;     // -1 for a leaf
  int32_t depth;
  uint32_t subtree;      // nodes in this subtree (bottom-up)
  uint32_t preorder;     // final index (top-down)
  T left_max, right_min;
  T left_min, right_max;  // the other two bounds of kd_tree_branch_double (topological metrics)
};

// one huge node of the current level
template <typename T>
struct HugeNode {
  uint32_t node;         // BFS id
  int32_t begin, end;
  int32_t sd;
  T split_val;
  uint32_t first_chunk;  // exclusive scan of the chunk counts over the level's huge list
  int32_t split;
  int32_t nl;            // points left of split_val
  int32_t m;             // misplaced pairs to exchange
  int32_t mode;          // 0 = partition by exchanges, 1 = nothing (more) to move
};

// what one chunk contributes to its node
template <typename T>
struct ChunkStat {
  int32_t lt;         // points with coord < split_val
  int32_t lt_prefix;  // exclusive prefix of lt inside the node (huge_resolve)
};

template <typename T>
struct BuildState {
  const T* raw;        // [n][sdim] packed row-major
  int32_t sdim;
  int32_t* idx;        // [n]
  int32_t* tmp;        // [n] scratch for the partition
  int32_t* tmp2;       // [n] second scratch list (nth_element emulation)
  BNode<T>* nodes;     // BFS table
  T* boxes;            // [cap][2*sdim]: cut box on the way down, tight box on the way up
  uint32_t* counters;  // [0] n_nodes  [1] n_big_next  [2] n_leaves  [3] n_huge_next  [4] chunks of this level  [5] largest leaf
  uint32_t* big_next;  // ids of next-level nodes that need a CTA
  uint32_t* huge_next; // ids of next-level nodes that need the chunked passes
  HugeNode<T>* huge;   // the current level's huge nodes
  ChunkStat<T>* chunks;
  int32_t rule, stop_kind, stop_value;
  int32_t huge_min;    // nodes with more points take the chunked passes (kHugeMin; PICO_B200_HUGE_MIN overrides)
};

// ------------------------------------------------------------------ cooperative groups
// G == 32: one warp. G > 32: the whole CTA (blockDim.x == G).
template <int G>
struct Grp {
  __device__ static int tid() { return threadIdx.x; }
  __device__ static void sync() { __syncthreads(); }
  __device__ static int sum(int v) {
    __shared__ int s_w[G / 32];
    __shared__ int s_tot;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int w = 0; w < G / 32; ++w) t += s_w[w];
      s_tot = t;
    }
    __syncthreads();
    return s_tot;
  }
  // exclusive rank of `flag` among the group's threads (thread order) and the total
  __device__ static int excl(int flag, int& total) {
    __shared__ int s_w[G / 32];
    const unsigned b = __ballot_sync(0xffffffffu, flag);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) s_w[w] = __popc(b);
    __syncthreads();
    int before = 0, tot = 0;
    for (int i = 0; i < G / 32; ++i) {
      const int c = s_w[i];
      if (i < w) before += c;
      tot += c;
    }
    total = tot;
    return before + __popc(b & ((1u << lane) - 1u));
  }
  // extreme of (value, position): larger value wins if MAX (smaller if !MAX), ties -> lower position
  template <typename T, bool MAX>
  __device__ static void arg_extreme(T& v, int& pos) {
    __shared__ T s_v[G / 32];
    __shared__ int s_p[G / 32];
    for (int o = 16; o > 0; o >>= 1) {
      const T ov = __shfl_xor_sync(0xffffffffu, v, o);
      const int op = __shfl_xor_sync(0xffffffffu, pos, o);
      const bool better = MAX ? (ov > v) : (ov < v);
      if (better || (ov == v && op < pos)) {
        v = ov;
        pos = op;
      }
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
      s_v[threadIdx.x >> 5] = v;
      s_p[threadIdx.x >> 5] = pos;
    }
    __syncthreads();
    v = s_v[0];
    pos = s_p[0];
    for (int w = 1; w < G / 32; ++w) {
      const T ov = s_v[w];
      const int op = s_p[w];
      const bool better = MAX ? (ov > v) : (ov < v);
      if (better || (ov == v && op < pos)) {
        v = ov;
        pos = op;
      }
    }
    __syncthreads();
  }
};

template <>
struct Grp<32> {
  __device__ static int tid() { return threadIdx.x & 31; }
  __device__ static void sync() { __syncwarp(); }
  __device__ static int sum(int v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  }
  __device__ static int excl(int flag, int& total) {
    const unsigned b = __ballot_sync(0xffffffffu, flag);
    total = __popc(b);
    return __popc(b & ((1u << (threadIdx.x & 31)) - 1u));
  }
  template <typename T, bool MAX>
  __device__ static void arg_extreme(T& v, int& pos) {
    for (int o = 16; o > 0; o >>= 1) {
      const T ov = __shfl_xor_sync(0xffffffffu, v, o);
      const int op = __shfl_xor_sync(0xffffffffu, pos, o);
      const bool better = MAX ? (ov > v) : (ov < v);
      if (better || (ov == v && op < pos)) {
        v = ov;
        pos = op;
      }
    }
  }
};

// ------------------------------------------------------------------ one node
template <typename T, int G>
__device__ void finish_leaf(const BuildState<T>& s, BNode<T>& nd, T* box) {
  const int tid = Grp<G>::tid();
  const int begin = nd.begin, end = nd.end, sdim = s.sdim;
  // "Keep the original box in case it is empty" (midpoint only, kd_tree_builder.hpp:363-371);
  // the other rules never produce an empty range.
  if (begin < end) {
    if (end - begin <= 64) {
      // few points: one thread per dimension, points visited in order
      for (int d = tid; d < sdim; d += G) {
        T mn = Limits<T>::max(), mx = -Limits<T>::max();
        for (int i = begin; i < end; ++i) {
          const T x = s.raw[(size_t)s.idx[i] * sdim + d];
          if (x < mn) mn = x;
          if (x > mx) mx = x;
        }
        box[d] = mn;
        box[sdim + d] = mx;
      }
    } else {
      for (int d = 0; d < sdim; ++d) {
        T mn = Limits<T>::max(), mx = -Limits<T>::max();
        int pn = 0, px = 0;
        for (int i = begin + tid; i < end; i += G) {
          const T x = s.raw[(size_t)s.idx[i] * sdim + d];
          if (x < mn) mn = x;
          if (x > mx) mx = x;
        }
        Grp<G>::template arg_extreme<T, false>(mn, pn);
        Grp<G>::template arg_extreme<T, true>(mx, px);
        if (tid == 0) {
          box[d] = mn;
          box[sdim + d] = mx;
        }
      }
    }
  }
  if (tid == 0) {
    nd.left = nd.right = -1;
    nd.split_dim = -1;
    nd.subtree = 1;
    atomicAdd(&s.counters[2], 1u);
    atomicMax(&s.counters[5], (uint32_t)(end - begin));  // largest leaf: sizes the staged leaf tiles
  }
}

// is_leaf of a node that does not exist yet (kd_tree_builder.hpp:410-425)
template <typename T>
__device__ __forceinline__ bool will_be_leaf(const BuildState<T>& s, int cnt, int depth) {
  return (s.stop_kind == PICO_B200_STOP_MAX_LEAF_SIZE) ? (cnt <= s.stop_value) : (depth == s.stop_value || cnt <= 1);
}

// Appends the two children of `node_id` to the BFS table, files them for the next level and
// writes their cut boxes (kd_tree_builder.hpp:379-383). Called by the whole group.
template <typename T, int G>
__device__ void emit_children(const BuildState<T>& s, uint32_t node_id, int sd, int split, T split_val) {
  BNode<T>& nd = s.nodes[node_id];
  const int tid = Grp<G>::tid();
  const int begin = nd.begin, end = nd.end, depth = nd.depth, sdim = s.sdim;
  const T* box = s.boxes + (size_t)node_id * 2 * sdim;
  if (tid == 0) {
    const uint32_t c = atomicAdd(&s.counters[0], 2u);
    nd.left = (int32_t)c;
    nd.right = (int32_t)c + 1;
    nd.split_dim = sd;
    BNode<T>& l = s.nodes[c];
    BNode<T>& r = s.nodes[c + 1];
    l.begin = begin;
    l.end = split;
    l.depth = depth + 1;
    r.begin = split;
    r.end = end;
    r.depth = depth + 1;
    l.left = l.right = r.left = r.right = -1;
    l.split_dim = r.split_dim = -1;
    const int cnts[2] = {split - begin, end - split};
    for (int k = 0; k < 2; ++k) {
      const bool huge = cnts[k] > s.huge_min && !will_be_leaf(s, cnts[k], depth + 1);
      if (huge)
        s.huge_next[atomicAdd(&s.counters[3], 1u)] = c + k;
      else if (cnts[k] > kWarpNodeMax)
        s.big_next[atomicAdd(&s.counters[1], 1u)] = c + k;
    }
  }
  Grp<G>::sync();
  const uint32_t cc = (uint32_t)nd.left;
  T* lb = s.boxes + (size_t)cc * 2 * sdim;
  T* rb = lb + 2 * sdim;
  for (int d = tid; d < 2 * sdim; d += G) {
    const T v = box[d];
    lb[d] = (d == sdim + sd) ? split_val : v;
    rb[d] = (d == sd) ? split_val : v;
  }
}

// ------------------------------------------------------------------ std::nth_element, exactly
// The reference calls std::nth_element in the slide cases of the sliding midpoint rule
// (kd_tree_builder.hpp:255-275) and for every node of the median rule (:167-173). Which index ends
// up where — beyond the n-th position itself — is decided by libstdc++'s introselect (GCC 13
// bits/stl_algo.h: median-of-three pivot, unguarded Hoare partition, insertion sort below four
// elements, heap select when the depth limit runs out), and the order of the indices inside a leaf
// decides which of two equidistant points a query reports. The emulation below leaves the index
// range in exactly that state. The Hoare partition is the only O(n) step and is data-parallel: the
// j-th element from the left that is >= pivot is exchanged with the j-th element from the right
// that is <= pivot for all j < m, where m counts the pairs whose left position is still smaller
// than the right one; the cut follows from the two position lists (tests/test_oracle_golden.py
// holds the sequential restatement this was checked against; tests/test_gpu_parity.py compares the
// resulting permutations with the reference's fixtures).
// CG: the index array is read past L1 (ld.cg) — for the multi-CTA flavour, where other CTAs exchange entries
// between two barriers of the same kernel.
template <typename T, bool CG = false>
struct NthCtx {
  using scalar = T;
  const T* col;  // coordinate `split_dim` of point p is col[p * sdim]
  int32_t* idx;
  int sdim;
  __device__ __forceinline__ int32_t id(int pos) const { return CG ? __ldcg(idx + pos) : idx[pos]; }
  __device__ __forceinline__ T at(int pos) const { return col[(size_t)id(pos) * sdim]; }
  __device__ __forceinline__ bool less(int a, int b) const { return at(a) < at(b); }  // positions
  __device__ __forceinline__ void swap(int a, int b) const {
    const int32_t t = id(a), u = id(b);
    idx[a] = u;
    idx[b] = t;
  }
};

// sequential pieces (one thread): std::__move_median_to_first, std::__insertion_sort, std::__heap_select
template <typename Ctx>
__device__ void seq_move_median_to_first(const Ctx& c, int result, int a, int b, int cc) {
  int pick;
  if (c.less(a, b)) {
    if (c.less(b, cc))
      pick = b;
    else if (c.less(a, cc))
      pick = cc;
    else
      pick = a;
  } else if (c.less(a, cc)) {
    pick = a;
  } else if (c.less(b, cc)) {
    pick = cc;
  } else {
    pick = b;
  }
  c.swap(result, pick);
}

template <typename T>
__device__ void seq_insertion_sort(const NthCtx<T>& c, int first, int last) {
  for (int i = first + 1; i < last; ++i) {
    const int32_t val = c.idx[i];
    const T v = c.col[(size_t)val * c.sdim];
    int j = i;
    if (v < c.at(first)) {
      for (; j > first; --j) c.idx[j] = c.idx[j - 1];
    } else {
      while (v < c.at(j - 1)) {
        c.idx[j] = c.idx[j - 1];
        --j;
      }
    }
    c.idx[j] = val;
  }
}

template <typename T>
__device__ void seq_adjust_heap(const NthCtx<T>& c, int first, int hole, int len, int32_t value) {
  const int top = hole;
  int child = hole;
  auto key = [&](int32_t id) { return c.col[(size_t)id * c.sdim]; };
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (c.less(first + child, first + child - 1)) child--;
    c.idx[first + hole] = c.idx[first + child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    c.idx[first + hole] = c.idx[first + child - 1];
    hole = child - 1;
  }
  int parent = (hole - 1) / 2;  // std::__push_heap
  while (hole > top && key(c.idx[first + parent]) < key(value)) {
    c.idx[first + hole] = c.idx[first + parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  c.idx[first + hole] = value;
}

template <typename T>
__device__ void seq_heap_select(const NthCtx<T>& c, int first, int middle, int last) {
  const int len = middle - first;
  if (len >= 2) {  // std::__make_heap
    for (int parent = (len - 2) / 2;; --parent) {
      seq_adjust_heap(c, first, parent, len, c.idx[first + parent]);
      if (parent == 0) break;
    }
  }
  for (int i = middle; i < last; ++i) {
    if (c.less(i, first)) {  // std::__pop_heap(first, middle, i)
      const int32_t value = c.idx[i];
      c.idx[i] = c.idx[first];
      seq_adjust_heap(c, first, 0, len, value);
    }
  }
}

// std::nth_element(first, nth, last) on positions of s.idx, by the whole group
// (depth_limit < 0: a fresh call; otherwise the introselect loop is resumed with that many levels left)
template <typename T, int G>
__device__ void group_nth_element(const BuildState<T>& s, int sd, int first, int nth, int last, int depth_limit = -1) {
  if (first == last || nth == last) return;
  const int tid = Grp<G>::tid();
  NthCtx<T> c{s.raw + sd, s.idx, s.sdim};
  int32_t* lst_l = s.tmp;   // positions of elements >= pivot, ascending, at [lo, lo + n_l)
  int32_t* lst_r = s.tmp2;  // positions of elements <= pivot, descending, at (hi - 1 - n_r, hi - 1]
  if (depth_limit < 0) depth_limit = 2 * (31 - __clz(last - first));
  while (last - first > 3) {
    if (depth_limit == 0) {
      if (tid == 0) {
        seq_heap_select(c, first, nth + 1, last);
        c.swap(first, nth);
      }
      Grp<G>::sync();
      return;
    }
    --depth_limit;
    if (tid == 0) seq_move_median_to_first(c, first, first + 1, first + (last - first) / 2, last - 1);
    Grp<G>::sync();
    const T pv = c.at(first);
    const int lo = first + 1, hi = last;
    int n_l = 0, n_r = 0;
    for (int base = lo; base < hi; base += G) {
      const int i = base + tid;
      const int f = (i < hi) && !(c.at(i) < pv);
      int tot;
      const int ex = Grp<G>::excl(f, tot);
      if (f) lst_l[lo + n_l + ex] = i;
      n_l += tot;
    }
    for (int base = 0; base < hi - lo; base += G) {
      const int i = hi - 1 - (base + tid);
      const int f = (i >= lo) && !(pv < c.at(i));
      int tot;
      const int ex = Grp<G>::excl(f, tot);
      if (f) lst_r[hi - 1 - (n_r + ex)] = i;
      n_r += tot;
    }
    Grp<G>::sync();
    const int pairs = n_l < n_r ? n_l : n_r;
    int cnt = 0;
    for (int j = tid; j < pairs; j += G) cnt += lst_l[lo + j] < lst_r[hi - 1 - j];
    const int m = Grp<G>::sum(cnt);
    for (int j = tid; j < m; j += G) c.swap(lst_l[lo + j], lst_r[hi - 1 - j]);
    int cut;
    if (m < n_l && (m == 0 || lst_l[lo + m] < lst_r[hi - 1 - (m - 1)]))
      cut = lst_l[lo + m];
    else
      cut = lst_r[hi - 1 - (m - 1)];
    Grp<G>::sync();
    if (cut <= nth)
      first = cut;
    else
      last = cut;
  }
  if (tid == 0) seq_insertion_sort(c, first, last);
  Grp<G>::sync();
}

template <typename T, int G>
__device__ void process_node(const BuildState<T>& s, uint32_t node_id) {
  BNode<T>& nd = s.nodes[node_id];
  const int tid = Grp<G>::tid();
  const int begin = nd.begin, end = nd.end, cnt = end - begin, depth = nd.depth;
  const int sdim = s.sdim;
  T* box = s.boxes + (size_t)node_id * 2 * sdim;

  // is_leaf, kd_tree_builder.hpp:410-425
  const bool leaf = (s.stop_kind == PICO_B200_STOP_MAX_LEAF_SIZE) ? (cnt <= s.stop_value)
                                                                  : (depth == s.stop_value || cnt <= 1);
  if (leaf) {
    finish_leaf<T, G>(s, nd, box);
    return;
  }

  // box_base::max_side, box.hpp:71-82 — strict '>' keeps the first longest side.
  int sd = 0;
  T max_delta = -Limits<T>::max();
  for (int d = 0; d < sdim; ++d) {
    const T delta = box[sdim + d] - box[d];
    if (delta > max_delta) {
      max_delta = delta;
      sd = d;
    }
  }
  const T box_min_sd = box[sd];
  const T* col = s.raw + sd;  // coordinate `sd` of point p is col[p * sdim]
  int32_t* idx = s.idx;
  int split;
  T split_val;

  {
    // kd_tree_builder.hpp:204 (midpoint: * 0.5) and :240 (sliding: / 2.0) — identical in
    // IEEE arithmetic, kept apart for the record.
    split_val = (s.rule == PICO_B200_RULE_MIDPOINT_MAX_SIDE) ? (max_delta * T(0.5) + box_min_sd)
                                                             : (max_delta / T(2.0) + box_min_sd);
    int nl = 0;
    for (int base = begin; base < end; base += G) {
      const int i = base + tid;
      nl += (i < end) && (col[(size_t)idx[i] * sdim] < split_val);
    }
    nl = Grp<G>::sum(nl);
    split = begin + nl;

    if (s.rule == PICO_B200_RULE_SLIDING_MIDPOINT_MAX_SIDE && (nl == cnt || nl == 0)) {
      // Nothing has moved (std::partition found every point on one side). One point slides to the
      // empty side: nth_element at the last / the second position, and the split value becomes the
      // coordinate found there (kd_tree_builder.hpp:255-275).
      split = (nl == cnt) ? end - 1 : begin + 1;
      group_nth_element<T, G>(s, sd, begin, split, end);
      split_val = col[(size_t)idx[split] * sdim];
    } else if (nl > 0 && nl < cnt) {
      // std::partition order: j-th misplaced from the left <-> j-th misplaced from the right
      int32_t* tmp = s.tmp;
      int m = 0;
      for (int base = begin; base < split; base += G) {
        const int i = base + tid;
        const int f = (i < split) && !(col[(size_t)idx[i] * sdim] < split_val);
        int tot;
        const int ex = Grp<G>::excl(f, tot);
        if (f) tmp[begin + m + ex] = i;
        m += tot;
      }
      if (m > 0) {
        int r = 0;
        for (int base = split; base < end; base += G) {
          const int i = base + tid;
          const int f = (i < end) && (col[(size_t)idx[i] * sdim] < split_val);
          int tot;
          const int ex = Grp<G>::excl(f, tot);
          if (f) tmp[end - 1 - (r + ex)] = i;
          r += tot;
        }
        Grp<G>::sync();
        for (int j = tid; j < m; j += G) {
          const int a = tmp[begin + j], b = tmp[end - m + j];
          const int32_t va = idx[a];
          idx[a] = idx[b];
          idx[b] = va;
        }
        Grp<G>::sync();
      }
    }
    // midpoint with nl == 0 or nl == cnt: one child is empty, nothing moves.
  }

  emit_children<T, G>(s, node_id, sd, split, split_val);
}

template <typename T>
__global__ void __launch_bounds__(256) split_level_warp(BuildState<T> s, uint32_t level_begin, uint32_t level_end) {
  const uint32_t node = level_begin + (blockIdx.x * blockDim.x + threadIdx.x) / 32;
  if (node >= level_end) return;
  if (s.nodes[node].end - s.nodes[node].begin > kWarpNodeMax) return;  // a CTA takes it
  process_node<T, 32>(s, node);
}

template <typename T>
__global__ void __launch_bounds__(kBigThreads) split_level_block(BuildState<T> s, const uint32_t* big_list) {
  process_node<T, kBigThreads>(s, big_list[blockIdx.x]);
}

// ------------------------------------------------------------------ huge nodes: chunked passes
// The exchanges std::partition makes are a function of the exclusive prefix count P(i) of
// "coord < split_val" flags alone: with nl = P(end) and split = begin + nl,
//   the j-th misplaced element of the left part  (i <  split, flag 0) has j = (i - begin) - P(i),
//   the j-th misplaced element from the right end (i >= split, flag 1) has j = nl - P(i) - 1,
// and element j of one list is exchanged with element j of the other. So a huge node needs one
// counting pass, one scan over its chunk counts, one pass that writes the two lists and one that
// exchanges — each spread over all SMs — instead of one CTA crawling over millions of indices.

// which huge node does chunk c belong to (largest h with first_chunk[h] <= c)
template <typename T>
__device__ __forceinline__ int huge_of_chunk(const HugeNode<T>* huge, int n_huge, uint32_t c) {
  int lo = 0, hi = n_huge - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (huge[mid].first_chunk <= c)
      lo = mid;
    else
      hi = mid - 1;
  }
  return lo;
}

// 1 CTA: split dimension / value of every huge node (box_base::max_side + kd_tree_builder.hpp:204,240)
// and the chunk layout of the level.
template <typename T>
__global__ void huge_plan(BuildState<T> s, const uint32_t* list, int n_huge) {
  const int sdim = s.sdim;
  for (int h = threadIdx.x; h < n_huge; h += blockDim.x) {
    const uint32_t id = list[h];
    const BNode<T>& nd = s.nodes[id];
    const T* box = s.boxes + (size_t)id * 2 * sdim;
    int sd = 0;
    T max_delta = -Limits<T>::max();
    for (int d = 0; d < sdim; ++d) {
      const T delta = box[sdim + d] - box[d];
      if (delta > max_delta) {
        max_delta = delta;
        sd = d;
      }
    }
    HugeNode<T>& hn = s.huge[h];
    hn.node = id;
    hn.begin = nd.begin;
    hn.end = nd.end;
    hn.sd = sd;
    hn.split_val = (s.rule == PICO_B200_RULE_MIDPOINT_MAX_SIDE) ? (max_delta * T(0.5) + box[sd])
                                                                : (max_delta / T(2.0) + box[sd]);
    hn.split = hn.nl = hn.m = 0;
    hn.mode = 1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t acc = 0;
    for (int h = 0; h < n_huge; ++h) {
      s.huge[h].first_chunk = acc;
      acc += (uint32_t)((s.huge[h].end - s.huge[h].begin + kChunk - 1) / kChunk);
    }
    s.counters[4] = acc;
  }
}

// grid = chunks of the level: points left of the split value, per chunk
template <typename T>
__global__ void __launch_bounds__(kChunkThreads) huge_count(BuildState<T> s, int n_huge) {
  const uint32_t c = blockIdx.x;
  if (c >= s.counters[4]) return;
  const HugeNode<T>& hn = s.huge[huge_of_chunk(s.huge, n_huge, c)];
  const int lo = hn.begin + (int)(c - hn.first_chunk) * kChunk;
  const int hi = min(lo + kChunk, hn.end);
  const T* col = s.raw + hn.sd;
  const T sv = hn.split_val;
  int lt = 0;
  for (int i = lo + threadIdx.x; i < hi; i += kChunkThreads) lt += col[(size_t)s.idx[i] * s.sdim] < sv;
  for (int o = 16; o > 0; o >>= 1) lt += __shfl_xor_sync(0xffffffffu, lt, o);
  __shared__ int s_lt[kChunkThreads / 32];
  if ((threadIdx.x & 31) == 0) s_lt[threadIdx.x >> 5] = lt;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kChunkThreads / 32; ++w) lt += s_lt[w];
    s.chunks[c].lt = lt;
    s.chunks[c].lt_prefix = 0;
  }
}

// one warp per huge node: scan of the chunk counts, slide fix-up, children
template <typename T>
__global__ void __launch_bounds__(32) huge_resolve(BuildState<T> s, int n_huge) {
  const int h = blockIdx.x;
  if (h >= n_huge) return;
  HugeNode<T>& hn = s.huge[h];
  const int lane = threadIdx.x;
  const int cnt = hn.end - hn.begin;
  const int n_chunks = (cnt + kChunk - 1) / kChunk;
  ChunkStat<T>* cs = s.chunks + hn.first_chunk;
  int carry = 0;
  for (int base = 0; base < n_chunks; base += 32) {
    const int c = base + lane;
    int v = c < n_chunks ? cs[c].lt : 0;
    int incl = v;
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (c < n_chunks) cs[c].lt_prefix = carry + incl - v;
    carry += __shfl_sync(0xffffffffu, incl, 31);
  }
  const int nl = carry;
  int split = hn.begin + nl;
  T split_val = hn.split_val;
  int mode = 0;
  if (s.rule == PICO_B200_RULE_SLIDING_MIDPOINT_MAX_SIDE && (nl == cnt || nl == 0)) {
    // slide (kd_tree_builder.hpp:255-275): std::nth_element at the last / second position, done by a
    // whole CTA in huge_slide
    split = (nl == cnt) ? hn.end - 1 : hn.begin + 1;
    mode = 2;
  } else if (nl == 0 || nl == cnt) {
    mode = 1;  // midpoint rule: one child is empty, nothing moves
  }
  if (lane == 0) {
    hn.nl = nl;
    hn.split = split;
    hn.mode = mode;
    hn.m = 0;
  }
  __syncwarp();
  if (mode != 2) emit_children<T, 32>(s, hn.node, hn.sd, split, split_val);
}

// one CTA per huge node that slides: exact nth_element, then the children
template <typename T>
__global__ void __launch_bounds__(kBigThreads) huge_slide(BuildState<T> s, int n_huge) {
  const HugeNode<T>& hn = s.huge[blockIdx.x];
  if (hn.mode != 2) return;
  group_nth_element<T, kBigThreads>(s, hn.sd, hn.begin, hn.split, hn.end);
  const T split_val = s.raw[(size_t)s.idx[hn.split] * s.sdim + hn.sd];
  __syncthreads();
  emit_children<T, kBigThreads>(s, hn.node, hn.sd, hn.split, split_val);
}

// grid = chunks: the two lists of misplaced positions (left list at tmp[begin + j], right list,
// counted from the right end, at tmp[end - 1 - j])
template <typename T>
__global__ void __launch_bounds__(kChunkThreads) huge_scatter(BuildState<T> s, int n_huge) {
  const uint32_t c = blockIdx.x;
  if (c >= s.counters[4]) return;
  HugeNode<T>& hn = s.huge[huge_of_chunk(s.huge, n_huge, c)];
  if (hn.mode != 0) return;
  const int lo = hn.begin + (int)(c - hn.first_chunk) * kChunk;
  const int hi = min(lo + kChunk, hn.end);
  const T* col = s.raw + hn.sd;
  const T sv = hn.split_val;
  const int begin = hn.begin, end = hn.end, split = hn.split, nl = hn.nl;
  __shared__ int s_w[kChunkThreads / 32];
  int base = s.chunks[c].lt_prefix;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int slab = lo; slab < hi; slab += kChunkThreads) {
    const int i = slab + threadIdx.x;
    const bool in = i < hi;
    const bool lt = in && (col[(size_t)s.idx[i] * s.sdim] < sv);
    const unsigned b = __ballot_sync(0xffffffffu, lt);
    __syncthreads();
    if (lane == 0) s_w[w] = __popc(b);
    __syncthreads();
    int before = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < kChunkThreads / 32; ++k) {
      const int v = s_w[k];
      if (k < w) before += v;
      tot += v;
    }
    const int P = base + before + __popc(b & ((1u << lane) - 1u));  // flags set in [begin, i)
    if (in) {
      if (i == split) hn.m = nl - P;
      if (i < split && !lt) s.tmp[begin + (i - begin) - P] = i;
      if (i >= split && lt) s.tmp[end - 1 - (nl - P - 1)] = i;
    }
    base += tot;
  }
}

template <typename T>
__global__ void __launch_bounds__(kChunkThreads) huge_swap(BuildState<T> s, int n_huge) {
  const uint32_t c = blockIdx.x;
  if (c >= s.counters[4]) return;
  const HugeNode<T>& hn = s.huge[huge_of_chunk(s.huge, n_huge, c)];
  if (hn.mode != 0) return;
  const int j0 = (int)(c - hn.first_chunk) * kChunk;
  const int j1 = min(j0 + kChunk, hn.m);
  for (int j = j0 + threadIdx.x; j < j1; j += kChunkThreads) {
    const int a = s.tmp[hn.begin + j], b = s.tmp[hn.end - 1 - j];
    const int32_t va = s.idx[a];
    s.idx[a] = s.idx[b];
    s.idx[b] = va;
  }
}

// ------------------------------------------------------------------ median rule
// nth_element at the middle (kd_tree_builder.hpp:153-176): the split position is fixed by the rule,
// the value is the coordinate that ends up there; group_nth_element leaves the indices exactly
// where libstdc++ would.
template <typename T, int G>
__device__ void median_node(const BuildState<T>& s, uint32_t node_id, bool selected = false) {
  BNode<T>& nd = s.nodes[node_id];
  const int tid = Grp<G>::tid();
  const int begin = nd.begin, end = nd.end, cnt = end - begin, depth = nd.depth;
  const int sdim = s.sdim;
  T* box = s.boxes + (size_t)node_id * 2 * sdim;
  const bool leaf = (s.stop_kind == PICO_B200_STOP_MAX_LEAF_SIZE) ? (cnt <= s.stop_value)
                                                                  : (depth == s.stop_value || cnt <= 1);
  if (leaf) {
    finish_leaf<T, G>(s, nd, box);
    return;
  }
  int sd = 0;
  T max_delta = -Limits<T>::max();
  for (int d = 0; d < sdim; ++d) {
    const T delta = box[sdim + d] - box[d];
    if (delta > max_delta) {
      max_delta = delta;
      sd = d;
    }
  }
  const T* col = s.raw + sd;
  int32_t* idx = s.idx;
  // std::nth_element at the middle (kd_tree_builder.hpp:165-175), emulated exactly
  const int split = begin + cnt / 2;
  if (!selected) group_nth_element<T, G>(s, sd, begin, split, end);
  const T split_val = col[(size_t)idx[split] * sdim];
  Grp<G>::sync();
  emit_children<T, G>(s, node_id, sd, split, split_val);
}

template <typename T>
__global__ void __launch_bounds__(256) median_level_warp(BuildState<T> s, uint32_t level_begin, uint32_t level_end) {
  const uint32_t node = level_begin + (blockIdx.x * blockDim.x + threadIdx.x) / 32;
  if (node >= level_end) return;
  if (s.nodes[node].end - s.nodes[node].begin > kWarpNodeMax) return;
  median_node<T, 32>(s, node);
}

template <typename T>
__global__ void __launch_bounds__(kBigThreads) median_level_block(BuildState<T> s, const uint32_t* big_list) {
  median_node<T, kBigThreads>(s, big_list[blockIdx.x]);
}

// Huge nodes of the median rule (the first ~8 levels of a multi-million point tree): one CTA per node would leave
// the GPU to 1, 2, 4 ... CTAs, each walking millions of indices (142-200 ms per 7.7 M points, profiles/r1). Here the
// CTAs of ONE co-resident grid (cooperative launch) are dealt out in equal groups over the level's huge nodes — the
// median rule keeps the nodes of a level within one point of each other — and a group runs the introselect
// iterations of its node together: the range is cut into one slice per CTA, every iteration takes
//   pivot (CTA 0)  |  count both flag kinds per slice  |  write the two position lists  |  exchange the pairs
// with a barrier of the group (a counter in global memory) after each step. The lists are what group_nth_element
// builds, so the exchanges and the cut are the same; once the range is short (or the depth limit is used up) CTA 0
// resumes the loop alone with the remaining depth limit. Everything the CTAs tell each other goes through L2
// (ld.cg / plain stores + __threadfence at the barrier).
__device__ __forceinline__ void group_barrier(unsigned* bar, unsigned& epoch, int ctas) {
  __syncthreads();
  if (ctas == 1) return;
  epoch += (unsigned)ctas;
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    while (*reinterpret_cast<volatile unsigned*>(bar) < epoch) {
    }
    __threadfence();
  }
  __syncthreads();
}

template <typename T>
__device__ void multi_nth_element(const BuildState<T>& s, int sd, int first, int nth, int last, int c, int ctas,
                                  unsigned* bar, unsigned& epoch, int32_t* cnt /* [2 * ctas] of this group */) {
  constexpr int B = kBigThreads;
  const int tid = threadIdx.x;
  NthCtx<T, true> ctx{s.raw + sd, s.idx, s.sdim};
  int32_t* lst_l = s.tmp;
  int32_t* lst_r = s.tmp2;
  const int multi_min = s.huge_min / 4;
  int depth_limit = 2 * (31 - __clz(last - first));
  while (last - first > multi_min && depth_limit > 0) {
    --depth_limit;
    if (c == 0 && tid == 0) seq_move_median_to_first(ctx, first, first + 1, first + (last - first) / 2, last - 1);
    group_barrier(bar, epoch, ctas);
    const T pv = ctx.at(first);
    const int lo = first + 1, hi = last;
    const int per = (hi - lo + ctas - 1) / ctas;
    const int my_lo = min(hi, lo + c * per), my_hi = min(hi, my_lo + per);
    // how many of each kind this slice holds
    {
      int nl = 0, nr = 0;
      for (int i = my_lo + tid; i < my_hi; i += B) {
        const T v = ctx.at(i);
        nl += !(v < pv);
        nr += !(pv < v);
      }
      nl = Grp<B>::sum(nl);
      nr = Grp<B>::sum(nr);
      if (tid == 0) {
        __stcg(cnt + c, nl);
        __stcg(cnt + ctas + c, nr);
      }
    }
    group_barrier(bar, epoch, ctas);
    int off_l = 0, off_r = 0, n_l = 0, n_r = 0;
    for (int o = 0; o < ctas; ++o) {
      const int a = __ldcg(cnt + o), b = __ldcg(cnt + ctas + o);
      if (o < c) off_l += a;
      if (o > c) off_r += b;
      n_l += a;
      n_r += b;
    }
    {
      int n = 0;
      for (int base = my_lo; base < my_hi; base += B) {
        const int i = base + tid;
        const int f = (i < my_hi) && !(ctx.at(i) < pv);
        int tot;
        const int ex = Grp<B>::excl(f, tot);
        if (f) lst_l[lo + off_l + n + ex] = i;
        n += tot;
      }
      n = 0;
      for (int base = 0; base < my_hi - my_lo; base += B) {
        const int i = my_hi - 1 - (base + tid);
        const int f = (i >= my_lo) && !(pv < ctx.at(i));
        int tot;
        const int ex = Grp<B>::excl(f, tot);
        if (f) lst_r[hi - 1 - (off_r + n + ex)] = i;
        n += tot;
      }
    }
    group_barrier(bar, epoch, ctas);
    // pairs j with lst_l[j] < lst_r[j]: the left positions ascend and the right ones descend, so they are a prefix
    const int pairs = n_l < n_r ? n_l : n_r;
    int m = 0;
    {
      int a = 0, b = pairs;  // first j in [0, pairs] that fails
      while (a < b) {
        const int j = (a + b) >> 1;
        if (__ldcg(lst_l + lo + j) < __ldcg(lst_r + hi - 1 - j))
          a = j + 1;
        else
          b = j;
      }
      m = a;
    }
    for (int j = c * B + tid; j < m; j += ctas * B) ctx.swap(__ldcg(lst_l + lo + j), __ldcg(lst_r + hi - 1 - j));
    int cut;
    if (m < n_l && (m == 0 || __ldcg(lst_l + lo + m) < __ldcg(lst_r + hi - 1 - (m - 1))))
      cut = __ldcg(lst_l + lo + m);
    else
      cut = __ldcg(lst_r + hi - 1 - (m - 1));
    group_barrier(bar, epoch, ctas);
    if (cut <= nth)
      first = cut;
    else
      last = cut;
  }
  if (c == 0) group_nth_element<T, B>(s, sd, first, nth, last, depth_limit);
}

// blockDim.x == kBigThreads; gridDim.x co-resident CTAs (cooperative launch); `ctas` CTAs per node
template <typename T>
__global__ void __launch_bounds__(kBigThreads) median_huge_level(BuildState<T> s, const uint32_t* huge_list,
                                                                 int n_huge, int ctas, unsigned* bars, int32_t* cnts) {
  const int groups = gridDim.x / ctas;
  const int g = blockIdx.x / ctas, c = blockIdx.x % ctas;
  if (g >= groups) return;
  unsigned epoch = 0;
  for (int h = g; h < n_huge; h += groups) {
    const uint32_t node_id = huge_list[h];
    const BNode<T>& nd = s.nodes[node_id];
    const int begin = nd.begin, end = nd.end, sdim = s.sdim;
    const T* box = s.boxes + (size_t)node_id * 2 * sdim;
    int sd = 0;
    T max_delta = -Limits<T>::max();
    for (int d = 0; d < sdim; ++d) {
      const T delta = box[sdim + d] - box[d];
      if (delta > max_delta) {
        max_delta = delta;
        sd = d;
      }
    }
    multi_nth_element<T>(s, sd, begin, begin + (end - begin) / 2, end, c, ctas, bars + g, epoch,
                         cnts + (size_t)g * 2 * ctas);
    if (c == 0) median_node<T, kBigThreads>(s, node_id, true);
  }
}

// ------------------------------------------------------------------ small kernels
template <typename T>
__global__ void root_box_kernel(const T* raw, size_t n, int sdim, T* partial /* [grid][2*sdim] */) {
  // one CTA computes min/max of every dimension over a grid-strided slice
  __shared__ T s_mn[32], s_mx[32];
  for (int d = 0; d < sdim; ++d) {
    T mn = Limits<T>::max(), mx = -Limits<T>::max();
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
      const T x = raw[i * sdim + d];
      if (x < mn) mn = x;
      if (x > mx) mx = x;
    }
    for (int o = 16; o > 0; o >>= 1) {
      const T a = __shfl_xor_sync(0xffffffffu, mn, o), b = __shfl_xor_sync(0xffffffffu, mx, o);
      if (a < mn) mn = a;
      if (b > mx) mx = b;
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
      s_mn[threadIdx.x >> 5] = mn;
      s_mx[threadIdx.x >> 5] = mx;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < (int)blockDim.x / 32; ++w) {
        if (s_mn[w] < mn) mn = s_mn[w];
        if (s_mx[w] > mx) mx = s_mx[w];
      }
      partial[(size_t)blockIdx.x * 2 * sdim + d] = mn;
      partial[(size_t)blockIdx.x * 2 * sdim + sdim + d] = mx;
    }
  }
}

template <typename T>
__global__ void root_box_final(const T* partial, int parts, int sdim, T* out) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= sdim) return;
  T mn = Limits<T>::max(), mx = -Limits<T>::max();
  for (int p = 0; p < parts; ++p) {
    const T a = partial[(size_t)p * 2 * sdim + d], b = partial[(size_t)p * 2 * sdim + sdim + d];
    if (a < mn) mn = a;
    if (b > mx) mx = b;
  }
  out[d] = mn;
  out[sdim + d] = mx;
}

__global__ void iota_kernel(int32_t* idx, size_t n) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) idx[i] = (int32_t)i;
}

// bottom-up over one level: tight boxes, stored bounds, subtree sizes (kd_tree_builder.hpp:388-392)
template <typename T>
__global__ void merge_level(BNode<T>* nodes, T* boxes, int sdim, uint32_t level_begin, uint32_t level_end) {
  const uint32_t node = level_begin + blockIdx.x * blockDim.x + threadIdx.x;
  if (node >= level_end) return;
  BNode<T>& nd = nodes[node];
  if (nd.split_dim < 0) return;
  const T* lb = boxes + (size_t)nd.left * 2 * sdim;
  const T* rb = boxes + (size_t)nd.right * 2 * sdim;
  T* box = boxes + (size_t)node * 2 * sdim;
  nd.left_max = lb[sdim + nd.split_dim];
  nd.right_min = rb[nd.split_dim];
  nd.left_min = lb[nd.split_dim];  // kd_tree_node_topological::set_branch, kd_tree_node.hpp:104-113
  nd.right_max = rb[sdim + nd.split_dim];
  for (int d = 0; d < sdim; ++d) {
    T mn = lb[d], mx = lb[sdim + d];
    if (rb[d] < mn) mn = rb[d];
    if (rb[sdim + d] > mx) mx = rb[sdim + d];
    box[d] = mn;
    box[sdim + d] = mx;
  }
  nd.subtree = 1 + nodes[nd.left].subtree + nodes[nd.right].subtree;
}

template <typename T>
__global__ void number_level(BNode<T>* nodes, uint32_t level_begin, uint32_t level_end) {
  const uint32_t node = level_begin + blockIdx.x * blockDim.x + threadIdx.x;
  if (node >= level_end) return;
  BNode<T>& nd = nodes[node];
  if (nd.split_dim < 0) return;
  nodes[nd.left].preorder = nd.preorder + 1;
  nodes[nd.right].preorder = nd.preorder + 1 + nodes[nd.left].subtree;
}

template <typename T>
__global__ void emit_nodes(const BNode<T>* nodes, uint32_t n_nodes, typename NodeOf<T>::type* out, T* outer,
                           uint2* spans) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_nodes) return;
  const BNode<T>& nd = nodes[i];
  typename NodeOf<T>::type o;
  memset(&o, 0, sizeof(o));
  if (nd.split_dim < 0) {
    o.a.begin_idx = nd.begin;
    o.b.end_idx = nd.end;
    o.right = PICO_B200_LEAF;
    o.split_dim = PICO_B200_LEAF;
  } else {
    o.a.left_max = nd.left_max;
    o.b.right_min = nd.right_min;
    o.right = nodes[nd.right].preorder;
    o.split_dim = (uint32_t)nd.split_dim;
  }
  out[nd.preorder] = o;
  if (spans) spans[nd.preorder] = make_uint2((uint32_t)nd.begin, (uint32_t)(nd.end - nd.begin));
  if (outer) {
    const bool leaf = nd.split_dim < 0;
    outer[2 * (size_t)nd.preorder] = leaf ? T(0) : nd.left_min;
    outer[2 * (size_t)nd.preorder + 1] = leaf ? T(0) : nd.right_max;
  }
}

// leaf-ordered point storage
template <typename T>
__global__ void pack_points4(const T* raw, const int32_t* idx, size_t n, int sdim, typename Vec4Of<T>::type* out) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t p = idx[i];
  typename Vec4Of<T>::type v;
  v.x = raw[(size_t)p * sdim];
  v.y = sdim > 1 ? raw[(size_t)p * sdim + 1] : T(0);
  v.z = sdim > 2 ? raw[(size_t)p * sdim + 2] : T(0);
  if (sizeof(T) == 4) {
    v.w = (T)__int_as_float(p);
  } else {
    v.w = (T)__longlong_as_double((long long)p);
  }
  out[i] = v;
}

template <typename T>
__global__ void pack_pointsN(const T* raw, const int32_t* idx, size_t n, int sdim, T* out) {
  // one warp per point row keeps both sides coalesced
  const size_t row = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) / 32;
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const T* src = raw + (size_t)idx[row] * sdim;
  T* dst = out + row * sdim;
  for (int d = lane; d < sdim; d += 32) dst[d] = src[d];
}

// Build-time scratch comes from the stream-ordered pool (kept warm by check_device's release
// threshold): a second build on the same device finds its ~1.5 GB of tables without a driver call.
struct DevBuf {
  void* p = nullptr;
  cudaStream_t st = nullptr;
  ~DevBuf() {
    if (p) cudaFreeAsync(p, st);
  }
  template <typename U>
  U* as() {
    return static_cast<U*>(p);
  }
};

int alloc(DevBuf& b, size_t bytes, cudaStream_t st) {
  b.st = st;
  PICO_CUDA(cudaMallocAsync(&b.p, bytes ? bytes : 16, st));
  return 0;
}

// Staging: rows (any stride; host memory, or device memory of this GPU — the kd_forest builds its trees over
// reflected copies that never leave the device) -> packed device rows.
template <typename T>
int stage_points(const T* h_pts, size_t n, size_t sdim, size_t stride, T* d_raw, cudaStream_t st) {
  if (stride == sdim)
    PICO_CUDA(cudaMemcpyAsync(d_raw, h_pts, n * sdim * sizeof(T), cudaMemcpyDefault, st));
  else
    PICO_CUDA(cudaMemcpy2DAsync(d_raw, sdim * sizeof(T), h_pts, stride * sizeof(T), sdim * sizeof(T), n,
                                cudaMemcpyDefault, st));
  return 0;
}

// The tree's own arrays are plain cudaMalloc blocks. Taking them from the stream-ordered pool like the workspaces
// makes a REBUILD marginally cheaper (9.4 against 9.8 ms) but leaves long-lived blocks inside the pool, and the
// multi-gigabyte result arrays of the ragged searches then cost 8-17 ms more per call to place (resident radius
// 21 -> 29-38 ms per call after other trees had been built and searched, profiles/r2/tree_pool.txt); the first
// full-size build of a process pays the pool's growth inside its timed region as well (13.5 -> 30-42 ms).
int tree_alloc(void** p, size_t bytes, cudaStream_t st) {
  static const bool pooled = [] {
    const char* e = getenv("PICO_B200_TREE_POOL");  // 1: from the stream-ordered pool (measurement hook)
    return e && atoi(e) == 1;
  }();
  if (pooled)
    PICO_CUDA(cudaMallocAsync(p, bytes ? bytes : 16, st));
  else
    PICO_CUDA(cudaMalloc(p, bytes ? bytes : 16));
  return 0;
}
template <typename P>
int tree_alloc(P** p, size_t bytes, cudaStream_t st) {
  return tree_alloc(reinterpret_cast<void**>(p), bytes, st);
}

template <typename T>
int finalize_storage(pico_b200_tree* t, const T* d_raw, cudaStream_t st) {
  const size_t n = t->n;
  const int sdim = (int)t->sdim;
  // (build_tree allocates the point storage before its timed region starts; the size does not depend on the tree)
  if (!t->d_pts) PICO_TRY(tree_alloc(&t->d_pts, t->pts_bytes() ? t->pts_bytes() : 16, st));
  if (t->packed()) {
    pack_points4<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_raw, t->d_indices, n, sdim,
                                                                  static_cast<typename Vec4Of<T>::type*>(t->d_pts));
  } else {
    pack_pointsN<T><<<(unsigned)((n * 32 + 255) / 256), 256, 0, st>>>(d_raw, t->d_indices, n, sdim,
                                                                      static_cast<T*>(t->d_pts));
  }
  PICO_CUDA(cudaGetLastError());
  t->device_bytes = t->pts_bytes() + t->n_nodes * t->node_size() + n * 4 + 2 * t->sdim * sizeof(T) + t->outer_bytes() +
                    t->spans_bytes();
  return 0;
}

}  // namespace

// Pinned host scratch of the calling thread, kept for the life of the thread: the per-level counters and the root
// box are read back through it. (A cudaMallocHost per build sat inside the timed region and cost 3-10 ms, now and
// then 30 — profiles/r2/build_timeline_median3.txt.)
void* pinned_scratch(size_t bytes) {
  struct Pin {
    void* p = nullptr;
    size_t cap = 0;
    ~Pin() {
      if (p) cudaFreeHost(p);
    }
  };
  thread_local Pin pin;
  if (bytes > pin.cap) {
    if (pin.p) cudaFreeHost(pin.p);
    pin.p = nullptr;
    pin.cap = 0;
    const size_t want = std::max<size_t>(bytes, 4096);
    if (cudaMallocHost(&pin.p, want) != cudaSuccess) {
      cudaGetLastError();
      pin.p = nullptr;
      return nullptr;
    }
    pin.cap = want;
  }
  return pin.p;
}

// root node, counters and empty lists in one launch (instead of five small copies out of pageable memory)
template <typename T>
__global__ void init_root(BuildState<T> s, int32_t n, const T* root_box, uint32_t* big, uint32_t* huge) {
  if (threadIdx.x == 0) {
    BNode<T> root;
    memset(&root, 0, sizeof(root));
    root.begin = 0;
    root.end = n;
    root.left = root.right = -1;
    root.split_dim = -1;
    root.depth = 0;
    root.preorder = 0;
    s.nodes[0] = root;
    for (int i = 0; i < 8; ++i) s.counters[i] = i == 0 ? 1u : 0u;
    big[0] = 0;
    huge[0] = 0;
  }
  for (int d = threadIdx.x; d < 2 * s.sdim; d += blockDim.x) s.boxes[d] = root_box[d];
}

// ------------------------------------------------------------------ host driver
template <typename T>
int build_tree(pico_b200_tree* t, const T* h_pts, size_t stride, int rule, int stop_kind, size_t stop_value,
               const T* bounds_min, const T* bounds_max) {
  const size_t n = t->n;
  const int sdim = (int)t->sdim;
  if (n >= (size_t)0x7fffffff) return fail(PICO_B200_ERR_UNSUPPORTED, "more than 2^31-2 points");
  // PICO_B200_BUILD_TIMELINE=1: host clock at the stages of the build and after every level's round trip, to stderr
  const bool timeline = getenv("PICO_B200_BUILD_TIMELINE") != nullptr;
  const auto tl0 = std::chrono::steady_clock::now();
  auto tl_ms = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tl0).count(); };
  cudaStream_t st;
  PICO_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  struct StreamGuard {
    cudaStream_t s;
    ~StreamGuard() { cudaStreamDestroy(s); }
  } guard{st};

  DevBuf raw, tmp, tmp2, nodes, boxes, counters, big_a, big_b, partial, huge_a, huge_b, huge_nodes, chunk_stats;
  PICO_TRY(alloc(raw, n * sdim * sizeof(T), st));
  PICO_TRY(stage_points(h_pts, n, sdim, stride, raw.as<T>(), st));
  PICO_TRY(tree_alloc(&t->d_indices, n * sizeof(int32_t), st));
  PICO_TRY(tree_alloc(&t->d_root_box, 2 * sdim * sizeof(T), st));
  // the leaf-ordered point storage: its size is known now, and a cudaMalloc of 120 MB inside the timed region made
  // build_ms swing between 10 and 25 ms
  PICO_TRY(tree_alloc(&t->d_pts, t->pts_bytes() ? t->pts_bytes() : 16, st));
  PICO_TRY(alloc(tmp, n * sizeof(int32_t), st));
  PICO_TRY(alloc(tmp2, n * sizeof(int32_t), st));

  // --- BFS node table
  size_t cap = 2 * n + 1024;
  PICO_TRY(alloc(nodes, cap * sizeof(BNode<T>), st));
  PICO_TRY(alloc(boxes, cap * 2 * sdim * sizeof(T), st));
  PICO_TRY(alloc(counters, 8 * sizeof(uint32_t), st));
  PICO_TRY(alloc(big_a, cap * sizeof(uint32_t), st));
  PICO_TRY(alloc(big_b, cap * sizeof(uint32_t), st));
  // huge nodes of one level are disjoint ranges of more than kHugeMin points each
  int huge_min = kHugeMin;
  if (const char* e = getenv("PICO_B200_HUGE_MIN")) huge_min = std::max(atoi(e), kWarpNodeMax);  // test hook
  const size_t huge_cap = n / (size_t)huge_min + 16;
  const size_t chunk_cap = n / kChunk + huge_cap + 16;
  PICO_TRY(alloc(huge_a, huge_cap * sizeof(uint32_t), st));
  PICO_TRY(alloc(huge_b, huge_cap * sizeof(uint32_t), st));
  PICO_TRY(alloc(huge_nodes, huge_cap * sizeof(HugeNode<T>), st));
  PICO_TRY(alloc(chunk_stats, chunk_cap * sizeof(ChunkStat<T>), st));

  // median rule: the grid of the cooperative kernel that takes the huge nodes, its barrier counters and slice counts
  int coop_grid = 0;
  DevBuf coop;
  if (rule == PICO_B200_RULE_MEDIAN_MAX_SIDE && n > (size_t)huge_min) {
    int per_sm = 0;
    PICO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, median_huge_level<T>, kBigThreads, 0));
    int can_cooperate = 0;
    cudaDeviceGetAttribute(&can_cooperate, cudaDevAttrCooperativeLaunch, t->device);
    coop_grid = can_cooperate ? per_sm * t->sm_count : 0;
    if (const char* e = getenv("PICO_B200_MEDIAN_COOP")) coop_grid = atoi(e) == 0 ? 0 : coop_grid;  // test hook
    PICO_TRY(alloc(coop, (size_t)std::max(coop_grid, 1) * 3 * sizeof(int32_t), st));
  }

  // build_ms covers the kernels and the per-level round trips, not the allocations above
  cudaEvent_t ev0, ev1;
  PICO_CUDA(cudaEventCreate(&ev0));
  PICO_CUDA(cudaEventCreate(&ev1));
  if (timeline) {
    fprintf(stderr, "points staged, workspaces allocated (host) at %.3f ms\n", tl_ms());
    cudaStreamSynchronize(st);
    fprintf(stderr, "... and done on the device at %.3f ms\n", tl_ms());
  }
  PICO_CUDA(cudaEventRecord(ev0, st));

  // --- root box
  T* d_root = static_cast<T*>(t->d_root_box);
  // [0, 64): the per-level counters; then the root box
  char* pinned = static_cast<char*>(pinned_scratch(64 + 2 * sdim * sizeof(T)));
  if (!pinned) return fail(PICO_B200_ERR_CUDA, "no pinned host memory for the build's read-backs");
  uint32_t* h_counters = reinterpret_cast<uint32_t*>(pinned);
  T* h_root = reinterpret_cast<T*>(pinned + 64);
  if (bounds_min && bounds_max) {
    // bbox.fit(min); bbox.fit(max)  kd_tree_builder.hpp:502-511
    for (int d = 0; d < sdim; ++d) {
      T mn = Limits<T>::max(), mx = -Limits<T>::max();
      const T xs[2] = {bounds_min[d], bounds_max[d]};
      for (T x : xs) {
        if (x < mn) mn = x;
        if (x > mx) mx = x;
      }
      h_root[d] = mn;
      h_root[sdim + d] = mx;
    }
    PICO_CUDA(cudaMemcpyAsync(d_root, h_root, 2 * sdim * sizeof(T), cudaMemcpyHostToDevice, st));
  } else {
    const int parts = (int)std::min<size_t>((n + 1023) / 1024, (size_t)t->sm_count * 4);
    PICO_TRY(alloc(partial, (size_t)parts * 2 * sdim * sizeof(T), st));
    root_box_kernel<T><<<parts, 1024, 0, st>>>(raw.as<T>(), n, sdim, partial.as<T>());
    root_box_final<T><<<(sdim + 127) / 128, 128, 0, st>>>(partial.as<T>(), parts, sdim, d_root);
    PICO_CUDA(cudaGetLastError());
    PICO_CUDA(cudaMemcpyAsync(h_root, d_root, 2 * sdim * sizeof(T), cudaMemcpyDeviceToHost, st));
    PICO_CUDA(cudaStreamSynchronize(st));
  }
  for (int d = 0; d < sdim && d < 4; ++d) {
    t->root_box_host[d] = (double)h_root[d];
    t->root_box_host[4 + d] = (double)h_root[sdim + d];
  }

  iota_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(t->d_indices, n);

  BuildState<T> s;
  s.raw = raw.as<T>();
  s.sdim = sdim;
  s.idx = t->d_indices;
  s.tmp = tmp.as<int32_t>();
  s.tmp2 = tmp2.as<int32_t>();
  s.nodes = nodes.as<BNode<T>>();
  s.boxes = boxes.as<T>();
  s.counters = counters.as<uint32_t>();
  s.huge = huge_nodes.as<HugeNode<T>>();
  s.chunks = chunk_stats.as<ChunkStat<T>>();
  s.huge_min = huge_min;
  s.rule = rule;
  s.stop_kind = stop_kind;
  s.stop_value = (int32_t)std::min<size_t>(stop_value, 0x7fffffff);

  init_root<T><<<1, 128, 0, st>>>(s, (int32_t)n, d_root, big_a.as<uint32_t>(), huge_a.as<uint32_t>());
  PICO_CUDA(cudaGetLastError());

  std::vector<uint32_t> level_start;  // BFS id of the first node of each level
  level_start.push_back(0);
  uint32_t level_begin = 0, level_end = 1;
  // the root: leaf / warp / CTA / chunked passes
  const bool root_leaf = (stop_kind == PICO_B200_STOP_MAX_LEAF_SIZE) ? (n <= (size_t)s.stop_value)
                                                                     : (s.stop_value == 0 || n <= 1);
  uint32_t n_huge = (n > (size_t)huge_min && !root_leaf) ? 1u : 0u;
  uint32_t n_big = (!n_huge && n > (size_t)kWarpNodeMax) ? 1u : 0u;
  uint32_t* big_cur = big_a.as<uint32_t>();
  uint32_t* big_nxt = big_b.as<uint32_t>();
  uint32_t* huge_cur = huge_a.as<uint32_t>();
  uint32_t* huge_nxt = huge_b.as<uint32_t>();

  if (timeline) fprintf(stderr, "root box, root node: level loop starts at %.3f ms\n", tl_ms());
  while (level_begin < level_end) {
    const uint32_t width = level_end - level_begin;
    if ((size_t)level_end + 2 * (size_t)width > cap) {
      // grow the tables (midpoint rule can create many empty leaves)
      const size_t ncap = std::max(cap * 2, (size_t)level_end + 2 * (size_t)width + 1024);
      DevBuf nn, nb, na, nbb;
      PICO_TRY(alloc(nn, ncap * sizeof(BNode<T>), st));
      PICO_TRY(alloc(nb, ncap * 2 * sdim * sizeof(T), st));
      PICO_TRY(alloc(na, ncap * sizeof(uint32_t), st));
      PICO_TRY(alloc(nbb, ncap * sizeof(uint32_t), st));
      PICO_CUDA(cudaMemcpyAsync(nn.p, nodes.p, cap * sizeof(BNode<T>), cudaMemcpyDeviceToDevice, st));
      PICO_CUDA(cudaMemcpyAsync(nb.p, boxes.p, cap * 2 * sdim * sizeof(T), cudaMemcpyDeviceToDevice, st));
      PICO_CUDA(cudaMemcpyAsync(na.p, big_cur, (size_t)n_big * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
      PICO_CUDA(cudaStreamSynchronize(st));
      std::swap(nodes.p, nn.p);
      std::swap(boxes.p, nb.p);
      std::swap(big_a.p, na.p);
      std::swap(big_b.p, nbb.p);
      big_cur = big_a.as<uint32_t>();
      big_nxt = big_b.as<uint32_t>();
      s.nodes = nodes.as<BNode<T>>();
      s.boxes = boxes.as<T>();
      cap = ncap;
    }
    s.big_next = big_nxt;
    s.huge_next = huge_nxt;
    PICO_CUDA(cudaMemsetAsync(s.counters + 1, 0, sizeof(uint32_t), st));
    PICO_CUDA(cudaMemsetAsync(s.counters + 3, 0, 2 * sizeof(uint32_t), st));
    if (n_huge && rule == PICO_B200_RULE_MEDIAN_MAX_SIDE && coop_grid > 0) {
      // equal groups of co-resident CTAs, one per huge node (as many nodes at a time as there are groups)
      const int ctas = std::max(1, coop_grid / (int)std::min<uint32_t>(n_huge, (uint32_t)coop_grid));
      PICO_CUDA(cudaMemsetAsync(coop.p, 0, (size_t)coop_grid * sizeof(unsigned), st));
      const uint32_t* list = huge_cur;
      int nh = (int)n_huge;
      unsigned* bars = coop.as<unsigned>();
      int32_t* cnts = coop.as<int32_t>() + coop_grid;
      void* args[] = {&s, &list, &nh, const_cast<int*>(&ctas), &bars, &cnts};
      if (cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(&median_huge_level<T>), dim3(coop_grid),
                                      dim3(kBigThreads), args, 0, st) != cudaSuccess) {
        // a device that cannot keep the grid resident (a partitioned or shared GPU): one CTA per huge node, the
        // slow but equivalent way
        cudaGetLastError();
        coop_grid = 0;
        median_level_block<T><<<n_huge, kBigThreads, 0, st>>>(s, huge_cur);
      }
    } else if (n_huge && rule == PICO_B200_RULE_MEDIAN_MAX_SIDE) {
      median_level_block<T><<<n_huge, kBigThreads, 0, st>>>(s, huge_cur);
    } else if (n_huge) {
      // every chunk kernel is launched over an upper bound of the level's chunk count
      const unsigned chunk_grid = (unsigned)std::min<size_t>(n / kChunk + n_huge + 1, chunk_cap);
      huge_plan<T><<<1, 256, 0, st>>>(s, huge_cur, (int)n_huge);
      huge_count<T><<<chunk_grid, kChunkThreads, 0, st>>>(s, (int)n_huge);
      huge_resolve<T><<<n_huge, 32, 0, st>>>(s, (int)n_huge);
      if (rule == PICO_B200_RULE_SLIDING_MIDPOINT_MAX_SIDE) huge_slide<T><<<n_huge, kBigThreads, 0, st>>>(s, (int)n_huge);
      huge_scatter<T><<<chunk_grid, kChunkThreads, 0, st>>>(s, (int)n_huge);
      huge_swap<T><<<chunk_grid, kChunkThreads, 0, st>>>(s, (int)n_huge);
    }
    const unsigned warp_blocks = (unsigned)(((size_t)width * 32 + 255) / 256);
    if (rule == PICO_B200_RULE_MEDIAN_MAX_SIDE) {
      median_level_warp<T><<<warp_blocks, 256, 0, st>>>(s, level_begin, level_end);
      if (n_big) median_level_block<T><<<n_big, kBigThreads, 0, st>>>(s, big_cur);
    } else {
      split_level_warp<T><<<warp_blocks, 256, 0, st>>>(s, level_begin, level_end);
      if (n_big) split_level_block<T><<<n_big, kBigThreads, 0, st>>>(s, big_cur);
    }
    PICO_CUDA(cudaGetLastError());
    PICO_CUDA(cudaMemcpyAsync(h_counters, s.counters, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    PICO_CUDA(cudaStreamSynchronize(st));
    if (timeline)
      fprintf(stderr, "level %zu: %u nodes (%u big, %u huge) done at %.3f ms\n", level_start.size() - 1, width, n_big,
              n_huge, tl_ms());
    level_begin = level_end;
    level_end = h_counters[0];
    n_big = h_counters[1];
    n_huge = h_counters[3];
    std::swap(big_cur, big_nxt);
    std::swap(huge_cur, huge_nxt);
    if (level_begin < level_end) level_start.push_back(level_begin);
    if ((size_t)level_end >= 0x7ffffff0u) return fail(PICO_B200_ERR_UNSUPPORTED, "node table overflow");
    if (level_start.size() > (1u << 20))
      return fail(PICO_B200_ERR_UNSUPPORTED, "tree deeper than 2^20 levels (degenerate input for this split rule)");
  }
  const uint32_t n_nodes = level_end;
  level_start.push_back(n_nodes);
  const size_t levels = level_start.size() - 1;
  t->n_nodes = n_nodes;
  t->n_leaves = h_counters[2];
  t->max_leaf_points = h_counters[5];
  t->height = levels - 1;

  // --- bottom-up (tight boxes, bounds, subtree sizes), then pre-order numbers
  for (size_t l = levels; l-- > 0;) {
    const uint32_t lb = level_start[l], le = level_start[l + 1];
    merge_level<T><<<(le - lb + 127) / 128, 128, 0, st>>>(s.nodes, s.boxes, sdim, lb, le);
  }
  for (size_t l = 0; l < levels; ++l) {
    const uint32_t lb = level_start[l], le = level_start[l + 1];
    number_level<T><<<(le - lb + 127) / 128, 128, 0, st>>>(s.nodes, lb, le);
  }
  PICO_TRY(tree_alloc(&t->d_nodes, (size_t)n_nodes * t->node_size(), st));
  if (t->outer_bytes()) PICO_TRY(tree_alloc(&t->d_outer, t->outer_bytes(), st));
  if (!t->packed()) PICO_TRY(tree_alloc(reinterpret_cast<void**>(&t->d_spans), t->spans_bytes(), st));
  emit_nodes<T><<<(n_nodes + 255) / 256, 256, 0, st>>>(s.nodes, n_nodes,
                                                        static_cast<typename NodeOf<T>::type*>(t->d_nodes),
                                                        static_cast<T*>(t->d_outer), t->d_spans);
  PICO_CUDA(cudaGetLastError());
  PICO_TRY(finalize_storage<T>(t, raw.as<T>(), st));
  PICO_TRY(build_fat_nodes(t, st));
  if (timeline) fprintf(stderr, "bottom-up, numbering, emit, point packing enqueued at %.3f ms\n", tl_ms());
  PICO_CUDA(cudaEventRecord(ev1, st));
  PICO_CUDA(cudaStreamSynchronize(st));
  if (timeline) fprintf(stderr, "build finished at %.3f ms\n", tl_ms());
  float ms = 0;
  PICO_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
  t->build_ms = ms;
  cudaEventDestroy(ev0);
  cudaEventDestroy(ev1);
  return 0;
}

// Upload of an existing tree (kd_tree::load path).
template <typename T>
int upload_tree(pico_b200_tree* t, const T* h_pts, size_t stride, const void* h_nodes, size_t n_nodes,
                const int32_t* indices, const T* root_box, const T* outer_bounds) {
  using NodeT = typename NodeOf<T>::type;
  const size_t n = t->n;
  const int sdim = (int)t->sdim;
  const NodeT* nodes = static_cast<const NodeT*>(h_nodes);
  // validate links and measure height / leaves on the host (cheap, once)
  size_t leaves = 0, height = 0, max_leaf = 0;
  {
    std::vector<std::pair<uint32_t, uint32_t>> stack;
    if (n_nodes == 0) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "tree has no nodes");
    stack.emplace_back(0u, 0u);
    size_t visited = 0;
    while (!stack.empty()) {
      auto [i, d] = stack.back();
      stack.pop_back();
      if (i >= n_nodes || ++visited > n_nodes) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "corrupt node links");
      height = std::max<size_t>(height, d);
      if (nodes[i].split_dim == PICO_B200_LEAF) {
        ++leaves;
        const int64_t b = nodes[i].a.begin_idx, e = nodes[i].b.end_idx;
        if (e > b) max_leaf = std::max<size_t>(max_leaf, (size_t)(e - b));
        if (b < 0 || e < b || (size_t)e > n) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "leaf range out of bounds");
      } else {
        if (nodes[i].split_dim >= (uint32_t)sdim) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "split_dim >= sdim");
        stack.emplace_back(nodes[i].right, d + 1);
        stack.emplace_back(i + 1, d + 1);
      }
    }
  }
  for (size_t i = 0; i < n; ++i)
    if (indices[i] < 0 || (size_t)indices[i] >= n) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "index out of range");
  t->n_nodes = n_nodes;
  t->n_leaves = leaves;
  t->max_leaf_points = max_leaf;
  t->height = height;

  cudaStream_t st;
  PICO_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  struct StreamGuard {
    cudaStream_t s;
    ~StreamGuard() { cudaStreamDestroy(s); }
  } guard{st};
  DevBuf raw;
  PICO_TRY(alloc(raw, n * sdim * sizeof(T), st));
  PICO_TRY(stage_points(h_pts, n, sdim, stride, raw.as<T>(), st));
  PICO_TRY(tree_alloc(&t->d_indices, n * sizeof(int32_t), st));
  PICO_TRY(tree_alloc(&t->d_root_box, 2 * sdim * sizeof(T), st));
  PICO_TRY(tree_alloc(&t->d_nodes, n_nodes * sizeof(NodeT), st));
  PICO_CUDA(cudaMemcpyAsync(t->d_indices, indices, n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  PICO_CUDA(cudaMemcpyAsync(t->d_root_box, root_box, 2 * sdim * sizeof(T), cudaMemcpyHostToDevice, st));
  PICO_CUDA(cudaMemcpyAsync(t->d_nodes, nodes, n_nodes * sizeof(NodeT), cudaMemcpyHostToDevice, st));
  if (!t->packed()) {
    // pre-order: children have larger ids than their parent, so one backward sweep suffices
    std::vector<uint2> spans(n_nodes);
    for (size_t i = n_nodes; i-- > 0;) {
      if (nodes[i].split_dim == PICO_B200_LEAF) {
        spans[i] = make_uint2((uint32_t)nodes[i].a.begin_idx, (uint32_t)(nodes[i].b.end_idx - nodes[i].a.begin_idx));
      } else {
        const uint2 l = spans[i + 1], r = spans[nodes[i].right];
        spans[i] = make_uint2(l.x, l.y + r.y);
      }
    }
    PICO_TRY(tree_alloc(reinterpret_cast<void**>(&t->d_spans), t->spans_bytes(), st));
    PICO_CUDA(cudaMemcpyAsync(t->d_spans, spans.data(), t->spans_bytes(), cudaMemcpyHostToDevice, st));
    PICO_CUDA(cudaStreamSynchronize(st));
  }
  if (t->topological()) {
    if (!outer_bounds) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "topological metric needs the outer bounds");
    PICO_TRY(tree_alloc(&t->d_outer, t->outer_bytes(), st));
    PICO_CUDA(cudaMemcpyAsync(t->d_outer, outer_bounds, t->outer_bytes(), cudaMemcpyHostToDevice, st));
  }
  for (int d = 0; d < sdim && d < 4; ++d) {
    t->root_box_host[d] = (double)root_box[d];
    t->root_box_host[4 + d] = (double)root_box[sdim + d];
  }
  PICO_TRY(finalize_storage<T>(t, raw.as<T>(), st));
  PICO_TRY(build_fat_nodes(t, st));
  PICO_CUDA(cudaStreamSynchronize(st));
  return 0;
}

template int build_tree<float>(pico_b200_tree*, const float*, size_t, int, int, size_t, const float*, const float*);
template int build_tree<double>(pico_b200_tree*, const double*, size_t, int, int, size_t, const double*,
                                const double*);
template int upload_tree<float>(pico_b200_tree*, const float*, size_t, const void*, size_t, const int32_t*,
                                const float*, const float*);
template int upload_tree<double>(pico_b200_tree*, const double*, size_t, const void*, size_t, const int32_t*,
                                 const double*, const double*);

}  // namespace pico
