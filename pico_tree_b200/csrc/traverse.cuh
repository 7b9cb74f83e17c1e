// traverse.cuh — device-side depth-first traversals of the flat tree.
//
// These are the CUDA counterparts of search_nearest_euclidean::search_nearest
// (internal/kd_tree_search.hpp:52-105) and search_box::operator()
// (internal/kd_tree_search.hpp:270-306). The recursion is unrolled onto an explicit stack;
// the arithmetic keeps the reference's operand order and is never contracted into FMAs
// (file compiled with --fmad=false and written with round-to-nearest intrinsics), so
// distances and pruning decisions are bit-identical to a portable x86-64 build of the
// reference.
#pragma once

#include "common.cuh"

namespace pico {

// ---------------------------------------------------------------- exact scalar ops
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float abs_t(float a) { return fabsf(a); }
__device__ __forceinline__ double abs_t(double a) { return fabs(a); }

__device__ __forceinline__ int index_of(const float4& p) { return __float_as_int(p.w); }
__device__ __forceinline__ int index_of(const double4& p) { return (int)__double_as_longlong(p.w); }

__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }
__device__ __forceinline__ double4 ldg4(const double4* p) {
  const double2 a = __ldg(reinterpret_cast<const double2*>(p));
  const double2 b = __ldg(reinterpret_cast<const double2*>(p) + 1);
  return make_double4(a.x, a.y, b.x, b.y);
}

// Node load: one 16-B (f32) / two 16-B (f64) read-only transactions.
struct NodeF32 {
  uint32_t a, b, right, split_dim;
};
__device__ __forceinline__ void load_node(const pico_b200_node_f32* nodes, uint32_t i, float& a, float& b,
                                          uint32_t& right, uint32_t& sd, int& lb, int& le) {
  const uint4 r = __ldg(reinterpret_cast<const uint4*>(nodes) + i);
  a = __uint_as_float(r.x);
  b = __uint_as_float(r.y);
  lb = (int)r.x;
  le = (int)r.y;
  right = r.z;
  sd = r.w;
}
__device__ __forceinline__ void load_node(const pico_b200_node_f64* nodes, uint32_t i, double& a, double& b,
                                          uint32_t& right, uint32_t& sd, int& lb, int& le) {
  const uint4* p = reinterpret_cast<const uint4*>(nodes) + 2 * (size_t)i;
  const uint4 r0 = __ldg(p);
  const uint4 r1 = __ldg(p + 1);
  a = __longlong_as_double((long long)(((unsigned long long)r0.y << 32) | r0.x));
  b = __longlong_as_double((long long)(((unsigned long long)r0.w << 32) | r0.z));
  lb = (int)r0.x;
  le = (int)r0.z;
  right = r1.x;
  sd = r1.y;
}

// metric(x): metric.hpp:93-96,119-122,147-150,176-179,210-213,240-243
template <typename T>
__device__ __forceinline__ T metric1(int metric, T x) {
  return (metric == PICO_B200_METRIC_L2_SQUARED || metric == PICO_B200_METRIC_SE2_SQUARED) ? mul_rn(x, x) : abs_t(x);
}

__device__ __forceinline__ bool is_topological(int metric) {
  return metric >= PICO_B200_METRIC_SO2 && metric <= PICO_B200_METRIC_CUSTOM_TOPOLOGICAL;
}

// s1_distance, distance.hpp:19-22: d = |x - y|; std::min(d, 1 - d)
template <typename T>
__device__ __forceinline__ T s1_distance(T x, T y) {
  const T d = abs_t(sub_rn(x, y));
  const T o = sub_rn(T(1.0), d);
  return o < d ? o : d;
}

// apply_dim_space: metric_so2 sees every dimension as the circle (metric.hpp:215-218),
// metric_se2_squared dimensions 0 and 1 as the line and the rest as the circle (:245-252).
__device__ __forceinline__ bool dim_is_s1(int metric, uint32_t dim) {
  return metric == PICO_B200_METRIC_SO2 || (metric == PICO_B200_METRIC_SE2_SQUARED && dim >= 2);
}

// One term folded into the running distance, j ascending (metric.hpp:36-51, :131-145, :158-174;
// so2 :203-208 looks at coordinate 0 only; se2_squared :229-238 sums two squared differences and
// adds the squared circle distance of the third coordinate).
template <typename T>
__device__ __forceinline__ T metric_fold(int metric, T d, T qj, T pj, int j) {
  const T t = sub_rn(qj, pj);
  switch (metric) {
    case PICO_B200_METRIC_L2_SQUARED:
      return add_rn(d, mul_rn(t, t));
    case PICO_B200_METRIC_L1:
      return add_rn(d, abs_t(t));
    case PICO_B200_METRIC_LPINF: {
      const T a = abs_t(t);
      return d < a ? a : d;
    }
    case PICO_B200_METRIC_LNINF: {
      const T a = abs_t(t);
      return a < d ? a : d;
    }
    case PICO_B200_METRIC_SO2:
      return j == 0 ? s1_distance(qj, pj) : d;
    default: {  // PICO_B200_METRIC_SE2_SQUARED
      if (j < 2) return add_rn(d, mul_rn(t, t));
      if (j > 2) return d;
      const T c = s1_distance(qj, pj);
      return add_rn(d, mul_rn(c, c));
    }
  }
}

// The first term of the fold (j = 0) on its own: the fold starts from metric_init (0, or the largest value for
// metric_lninf), and 0 + x, max(0, |x|) and min(largest, |x|) are x / |x| exactly — one addition per point saved.
template <typename T>
__device__ __forceinline__ T metric_first(int metric, T q0, T p0) {
  const T t = sub_rn(q0, p0);
  switch (metric) {
    case PICO_B200_METRIC_L2_SQUARED:
    case PICO_B200_METRIC_SE2_SQUARED:
      return mul_rn(t, t);
    case PICO_B200_METRIC_SO2:
      return s1_distance(q0, p0);
    default:
      return abs_t(t);
  }
}

// search_nearest_topological::box_distance (kd_tree_search.hpp:205-229): distance of v to the
// segment [mn, mx] on the line (segment.hpp:34-42) or on the circle (:76-100), then metric(d).
template <typename T>
__device__ __forceinline__ T topo_box_distance(int metric, T mn, T mx, T v, uint32_t dim) {
  T d = T(0);
  if (!dim_is_s1(metric, dim)) {
    if (v < mn)
      d = sub_rn(mn, v);
    else if (v > mx)
      d = sub_rn(v, mx);
  } else {
    const T a = s1_distance(v, mn), b = s1_distance(v, mx);
    const T m = b < a ? b : a;
    if (mn <= mx) {
      if (v < mn || v > mx) d = m;
    } else {
      if (!(v < mx || v > mn)) d = m;
    }
  }
  return metric1(metric, d);
}

// {left_min, right_max} of node i (kd_tree_branch_double's two extra bounds, kd_tree_node.hpp:52-59)
__device__ __forceinline__ void load_outer(const float* outer, uint32_t i, float& lmin, float& rmax) {
  const float2 r = __ldg(reinterpret_cast<const float2*>(outer) + i);
  lmin = r.x;
  rmax = r.y;
}
__device__ __forceinline__ void load_outer(const double* outer, uint32_t i, double& lmin, double& rmax) {
  const double2 r = __ldg(reinterpret_cast<const double2*>(outer) + i);
  lmin = r.x;
  rmax = r.y;
}

// Which child is nearer, and the box offset of the other one on split_dim.
//   euclidean   (kd_tree_search.hpp:76-88): left iff (left_max + right_min - v - v) > 0, evaluated left
//               to right; offset = metric(bound - v)
//   topological (kd_tree_search.hpp:166-186): left iff d(left box) < d(right box); offset = the larger one
template <typename T>
__device__ __forceinline__ void branch_choice(int metric, const T* __restrict__ outer, uint32_t node, T a, T b, T v,
                                              uint32_t sd, bool& go_left, T& new_off) {
  if (is_topological(metric)) {
    T lmin, rmax;
    load_outer(outer, node, lmin, rmax);
    const T d1 = topo_box_distance(metric, lmin, a, v, sd);
    const T d2 = topo_box_distance(metric, b, rmax, v, sd);
    go_left = d1 < d2;
    new_off = go_left ? d2 : d1;
  } else {
    go_left = sub_rn(sub_rn(add_rn(a, b), v), v) > T(0);
    new_off = metric1(metric, sub_rn(go_left ? b : a, v));
  }
}
template <typename T>
__device__ __forceinline__ T metric_init(int metric) {
  return metric == PICO_B200_METRIC_LNINF ? Limits<T>::max() : T(0);
}

// ---------------------------------------------------------------- thread-per-query, sdim <= 3
// Stack entry = far child + its box distance + a snapshot of the per-dimension offsets
// (node_box_offset_) that hold while the far subtree is visited. Taking a snapshot instead
// of the reference's set/restore pair is equivalent: every pop installs the full state.
template <typename T, int DIM, int DEPTH>
struct LocalStack {
  uint32_t node[DEPTH];
  T dist[DEPTH];
  T off[DIM][DEPTH];
  __device__ __forceinline__ void push(int sp, uint32_t n, T d, const T (&o)[DIM]) {
    node[sp] = n;
    dist[sp] = d;
#pragma unroll
    for (int j = 0; j < DIM; ++j) off[j][sp] = o[j];
  }
  __device__ __forceinline__ void pop(int sp, uint32_t& n, T& d, T (&o)[DIM]) const {
    n = node[sp];
    d = dist[sp];
#pragma unroll
    for (int j = 0; j < DIM; ++j) o[j] = off[j][sp];
  }
};

// Same interface over a global workspace (trees deeper than the local stack). Entries of
// one thread are `stride` words apart so that a warp's accesses coalesce.
template <typename T, int DIM>
struct GlobalStack {
  uint32_t* node;
  T* dist;
  T* off;  // [DIM][depth][stride]
  size_t stride, depth;
  __device__ __forceinline__ void push(int sp, uint32_t n, T d, const T (&o)[DIM]) {
    node[(size_t)sp * stride] = n;
    dist[(size_t)sp * stride] = d;
#pragma unroll
    for (int j = 0; j < DIM; ++j) off[((size_t)j * depth + sp) * stride] = o[j];
  }
  __device__ __forceinline__ void pop(int sp, uint32_t& n, T& d, T (&o)[DIM]) const {
    n = node[(size_t)sp * stride];
    d = dist[(size_t)sp * stride];
#pragma unroll
    for (int j = 0; j < DIM; ++j) o[j] = off[((size_t)j * depth + sp) * stride];
  }
};

// FAST = metric_l2_squared + exact visitors, resolved at compile time.
//
// PRIME (used by the single-neighbour visitors): the first root-to-leaf descent is walked
// once WITHOUT recording the far children, the first leaf is scanned, and the traversal
// then restarts from the root with max() already finite. The reference pushes (recursion
// frames) every far child of that first descent while max() is still FLT_MAX, and rejects
// almost all of them later; restarting lets the `max() >= far_dist` test drop them before
// they ever touch the stack. The restart re-reads ~25 nodes that are warp-coherent L1 hits,
// and saves ~25 x 20 B of per-thread stack stores per query that were the kernel's main
// DRAM traffic (profiles/r1/knn1_v1_summary.txt: 3.4 GB written per launch).
// Equivalence: the second walk makes the same near/far choices, visits the far children in
// the same (deepest-first) order, tests them against a max() that is never larger than
// the one the reference sees, and skips the already scanned first leaf.
//
// PRIME == 2 (exact search, k > 1): a k-list stays "infinite" until k points were seen, i.e. for
// the whole first descent and the first leaf or two — the same ~25 unconditional frames per
// query (3.9 GB of stack traffic per launch at k = 16, a third of the kernel's stall samples).
// Here the first descent (again without frames) only locates the first leaf; the largest
// distance among k consecutive stored points around it (leaf order keeps them spatially close)
// is an upper bound B of the final k-th distance, and the real traversal starts from the root
// pruning with min(max(), B). Nothing is inserted out of order. A node dropped by B but visited
// by the reference holds only points farther than B >= the final k-th distance: none of them
// survives in the reference's list, and removing such points from the visit sequence changes
// neither the final members nor their order (insert_sorted is stable). Not used for the
// approximate visitors, whose lists are not the true k nearest, nor for metric_lpinf / metric_lninf,
// nor for trees deeper than the local stack (the rounding margin of the bound assumes depth < 64).
constexpr int kPrimeNone = 0, kPrimeFirstLeaf = 1, kPrimeBound = 2;

template <typename T, int DIM, bool FAST, int PRIME, typename Stack, typename Visitor>
__device__ __forceinline__ void traverse_packed(const typename NodeOf<T>::type* __restrict__ nodes,
                                                const typename Vec4Of<T>::type* __restrict__ pts4,
                                                const T* __restrict__ outer, const T (&q)[DIM], int metric_rt,
                                                bool approx_rt, T e_inv, Stack& stack, Visitor& vis,
                                                int n_points = 0, int k = 0) {
  const int metric = FAST ? (int)PICO_B200_METRIC_L2_SQUARED : metric_rt;
  const bool approx = FAST ? false : approx_rt;
  T bound = Limits<T>::max();
  T off[DIM];
#pragma unroll
  for (int j = 0; j < DIM; ++j) off[j] = T(0);
  uint32_t node = 0;
  T node_dist = T(0);
  int sp = 0;
  uint32_t primed_leaf = 0xFFFFFFFEu;
  // (the bound needs box distances that are true lower bounds of the point distances: the reference sums
  // per-dimension offsets for every metric, which over-estimates under metric_lpinf / metric_lninf — there
  // its pruning is part of the result and must be reproduced exactly)
  const bool bound_ok = metric != PICO_B200_METRIC_LPINF && metric != PICO_B200_METRIC_LNINF;
  if (PRIME == kPrimeFirstLeaf || (PRIME == kPrimeBound && bound_ok && !approx && k > 1 && n_points >= k)) {
    T a, b;
    uint32_t right, sd;
    int lb, le;
    load_node(nodes, node, a, b, right, sd, lb, le);
    while (sd != PICO_B200_LEAF) {
      T v = q[0];
#pragma unroll
      for (int j = 1; j < DIM; ++j)
        if (sd == (uint32_t)j) v = q[j];
      bool go_left;
      T unused;
      branch_choice(metric, outer, node, a, b, v, sd, go_left, unused);
      node = go_left ? node + 1 : right;
      load_node(nodes, node, a, b, right, sd, lb, le);
    }
    if (PRIME == kPrimeFirstLeaf) {
      for (int i = lb; i < le; ++i) {
        const typename Vec4Of<T>::type p = ldg4(pts4 + i);
        T d = metric_first(metric, q[0], p.x);
        if (DIM > 1) d = metric_fold(metric, d, q[DIM > 1 ? 1 : 0], p.y, 1);
        if (DIM > 2) d = metric_fold(metric, d, q[DIM > 2 ? 2 : 0], p.z, 2);
        if (approx) d = mul_rn(d, e_inv);
        vis.visit(index_of(p), d);
      }
      primed_leaf = node;
    } else {
      // k consecutive stored points centred on the first leaf
      int s = lb - (k - (le - lb)) / 2;
      s = s < 0 ? 0 : (s > n_points - k ? n_points - k : s);
      T far = T(0);
      for (int i = s; i < s + k; ++i) {
        const typename Vec4Of<T>::type p = ldg4(pts4 + i);
        T d = metric_first(metric, q[0], p.x);
        if (DIM > 1) d = metric_fold(metric, d, q[DIM > 1 ? 1 : 0], p.y, 1);
        if (DIM > 2) d = metric_fold(metric, d, q[DIM > 2 ? 2 : 0], p.z, 2);
        far = d > far ? d : far;
      }
      // far_dist values are sums built with one rounding per tree level and can exceed the exact
      // box distance by a few ulps; a point that IS the bound (the window may reach into the far
      // node) must not lose its own node to that: widen by 1024 eps >> 2 * depth * eps (depth < 64)
      bound = add_rn(far, mul_rn(far, sizeof(T) == 4 ? T(1.2207031e-4) : T(2.2737368e-13)));
    }
    node = 0;
  }
  // max() of the visitor, capped by the primed bound
  auto reach = [&]() -> T {
    const T m = vis.max();
    return (PRIME == kPrimeBound && bound < m) ? bound : m;
  };
  for (;;) {
    // ---- descend to a leaf (kd_tree_search.hpp:60-88)
    T a, b;
    uint32_t right, sd;
    int lb, le;
    load_node(nodes, node, a, b, right, sd, lb, le);
    while (sd != PICO_B200_LEAF) {
      T v = q[0], old = off[0];
#pragma unroll
      for (int j = 1; j < DIM; ++j) {
        if (sd == (uint32_t)j) {
          v = q[j];
          old = off[j];
        }
      }
      bool go_left;
      T new_off;
      branch_choice(metric, outer, node, a, b, v, sd, go_left, new_off);
      const uint32_t far = go_left ? right : node + 1;
      node = go_left ? node + 1 : right;
      // node_box_distance - old_offset + new_offset (kd_tree_search.hpp:93-94)
      const T far_dist = add_rn(sub_rn(node_dist, old), new_off);
      // The reference tests `visitor.max() >= dist` after the near subtree; max() never
      // grows, so a far child that already fails now can be dropped without a push.
      if (reach() >= far_dist) {
        T snap[DIM];
#pragma unroll
        for (int j = 0; j < DIM; ++j) snap[j] = (sd == (uint32_t)j) ? new_off : off[j];
        stack.push(sp, far, far_dist, snap);
        ++sp;
      }
      load_node(nodes, node, a, b, right, sd, lb, le);
    }
    // ---- leaf scan (kd_tree_search.hpp:54-59): contiguous Vec4 records, index in .w
    if (PRIME == kPrimeFirstLeaf && node == primed_leaf) le = lb;
    for (int i = lb; i < le; ++i) {
      const typename Vec4Of<T>::type p = ldg4(pts4 + i);
      T d = metric_first(metric, q[0], p.x);
      if (DIM > 1) d = metric_fold(metric, d, q[DIM > 1 ? 1 : 0], p.y, 1);
      if (DIM > 2) d = metric_fold(metric, d, q[DIM > 2 ? 2 : 0], p.z, 2);
      if (approx) d = mul_rn(d, e_inv);
      vis.visit(index_of(p), d);
    }
    // ---- next pending far child (kd_tree_search.hpp:99-103)
    bool found = false;
    while (sp > 0) {
      --sp;
      T d;
      uint32_t n;
      T snap[DIM];
      stack.pop(sp, n, d, snap);
      if (reach() >= d) {
        node = n;
        node_dist = d;
#pragma unroll
        for (int j = 0; j < DIM; ++j) off[j] = snap[j];
        found = true;
        break;
      }
    }
    if (!found) return;
  }
}

// Eight words to a 32-byte aligned address in ONE store (sm_100: STG.E.ENL2.256): a whole sector arrives at L2 at
// once, which neither reads it from DRAM first nor writes it back in pieces.
__device__ __forceinline__ void store_sector(void* p, int4 lo, int4 hi) {
#if defined(__CUDA_ARCH__)
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(lo.x), "r"(lo.y), "r"(lo.z),
               "r"(lo.w), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w)
               : "memory");
#else
  reinterpret_cast<int4*>(p)[0] = lo;
  reinterpret_cast<int4*>(p)[1] = hi;
#endif
}

// One neighbour record. {int, float} has an alignment of four, so the compiler writes it as two 4-byte stores; rows that
// sit on an 8-byte boundary (every buffer a CUDA or host allocator hands out) take one 8-byte store instead — half
// the store instructions and no half-written records on their way through L2.
template <typename T>
__device__ __forceinline__ void store_neighbor(Neighbor<T>* p, int idx, T d) {
  if constexpr (sizeof(Neighbor<T>) == 8) {
    if ((reinterpret_cast<uintptr_t>(p) & 7) == 0) {
      *reinterpret_cast<int2*>(p) = make_int2(idx, __float_as_int(d));
      return;
    }
  }
  p->index = idx;
  p->distance = d;
}

// ---------------------------------------------------------------- visitors (thread-local)
// search_nn, search_visitor.hpp:41-65 (+ approximate :164-191; scaling done by the caller)
template <typename T>
struct VisitNn {
  T best = Limits<T>::max();
  int idx = -1;
  __device__ __forceinline__ T max() const { return best; }
  __device__ __forceinline__ void visit(int i, T d) {
    if (best > d) {
      best = d;
      idx = i;
    }
  }
};

// search_knn + insert_sorted, search_visitor.hpp:20-38,82-123. The k-list lives in the LAST k of
// KMAX register slots; the slots in front of it hold a negative sentinel that no distance can
// displace, so the insertion network is the same for every k <= KMAX and max() is always slot
// KMAX-1 — a compile-time index. (Reading slot k-1 with a run-time k makes the compiler index the
// array dynamically, which moves it to local memory: 6.4 GB of DRAM writes per launch at k = 16,
// profiles/r1/knn16_v3_summary.txt.) Equal distances keep the earlier-visited neighbour first.
template <typename T, int KMAX>
struct VisitKnn {
  T d[KMAX];
  int id[KMAX];
  int k;
  __device__ __forceinline__ void init(int k_) {
    k = k_;
#pragma unroll
    for (int i = 0; i < KMAX; ++i) {
      d[i] = (i < KMAX - k_) ? T(-1) : Limits<T>::max();
      id[i] = -1;
    }
  }
  __device__ __forceinline__ T max() const { return d[KMAX - 1]; }
  // The network runs from the tail in segments of four slots and stops at the first segment boundary whose lower
  // neighbour is not greater than x: everything below stays as it is. Candidates accepted late in a search land
  // near the tail of the list, so most insertions touch one or two segments instead of all KMAX slots.
  __device__ __forceinline__ void visit(int i_new, T x) {
    if (!(d[KMAX - 1] > x)) return;
    constexpr int SEG = KMAX >= 8 ? 4 : KMAX;
#pragma unroll
    for (int hi = KMAX - 1; hi >= 0; hi -= SEG) {
#pragma unroll
      for (int i = hi; i > hi - SEG; --i) {
        if (i >= 1) {
          const bool shift = d[i - 1] > x;
          const bool here = !shift && (d[i] > x);
          const T nd = shift ? d[i - 1] : (here ? x : d[i]);
          const int ni = shift ? id[i - 1] : (here ? i_new : id[i]);
          d[i] = nd;
          id[i] = ni;
        } else if (d[0] > x) {
          d[0] = x;
          id[0] = i_new;
        }
      }
      if (hi - SEG >= 0 && !(d[hi - SEG] > x)) return;
    }
  }
  // row = k neighbour records, ascending
  __device__ __forceinline__ void store(Neighbor<T>* row) const {
    if constexpr (sizeof(Neighbor<T>) == 8 && KMAX % 4 == 0) {
      // a full list of 8-byte records at a 32-byte boundary leaves as KMAX / 4 whole sectors
      if (k == KMAX && (reinterpret_cast<uintptr_t>(row) & 31) == 0) {
#pragma unroll
        for (int i = 0; i < KMAX; i += 4)
          store_sector(row + i, make_int4(id[i], __float_as_int(d[i]), id[i + 1], __float_as_int(d[i + 1])),
                       make_int4(id[i + 2], __float_as_int(d[i + 2]), id[i + 3], __float_as_int(d[i + 3])));
        return;
      }
    }
    if constexpr (sizeof(Neighbor<T>) == 8 && KMAX % 2 == 0) {
      // ... at a 16-byte boundary as KMAX / 2 16-byte stores
      if (k == KMAX && (reinterpret_cast<uintptr_t>(row) & 15) == 0) {
#pragma unroll
        for (int i = 0; i < KMAX; i += 2)
          reinterpret_cast<int4*>(row)[i / 2] = make_int4(id[i], __float_as_int(d[i]), id[i + 1], __float_as_int(d[i + 1]));
        return;
      }
    }
    Neighbor<T>* shifted = row - (KMAX - k);
#pragma unroll
    for (int i = 0; i < KMAX; ++i) {
      if (i >= KMAX - k) {
        store_neighbor(shifted + i, id[i], d[i]);
      }
    }
  }
};

// search_radius, search_visitor.hpp:126-156: max() is the constant radius.
template <typename T>
struct VisitRadiusCount {
  T radius;
  uint32_t count = 0;
  __device__ __forceinline__ T max() const { return radius; }
  __device__ __forceinline__ void visit(int, T d) { count += (radius > d); }
};
// Fill pass. Every thread writes its query's hits one behind the other, so a store of one 8-byte neighbour touches a
// quarter of a 32-byte sector, and 32 lanes touch 32 different sectors: the fill pass of cfg3 moved 3.8 GB to DRAM and
// 2.1 GB from it for 1.7 GB of hits and took 2.2x the count pass (profiles/r2/late_search_ncu_summary.txt). For 8-byte
// neighbours the hits of one SECTOR of the result are therefore collected in registers and leave as two 16-byte stores
// when the sector is complete; only the first and the last sector of a query, which it may share with its
// neighbours in the array, are written hit by hit (begin / finish). The order of the hits is untouched.
template <typename T>
struct VisitRadiusFill {
  T radius;
  Neighbor<T>* out;
  int2 h0, h1, h2, h3;  // the hits of the current sector (8-byte neighbours only)
  int first;            // slot of the first hit staged for it, 4 = none
  static constexpr bool kStaged = sizeof(Neighbor<T>) == 8;
  __device__ __forceinline__ void begin(Neighbor<T>* o) {
    out = o;
    first = 4;
  }
  __device__ __forceinline__ T max() const { return radius; }
  __device__ __forceinline__ int slot_of(const Neighbor<T>* p) const {
    return (int)((reinterpret_cast<uintptr_t>(p) >> 3) & 3);
  }
  __device__ __forceinline__ void visit(int i, T d) {
    if (!(radius > d)) return;
    if constexpr (kStaged) {
      const int slot = slot_of(out);
      const int2 v = make_int2(i, __float_as_int(d));
      if (slot == 0)
        h0 = v;
      else if (slot == 1)
        h1 = v;
      else if (slot == 2)
        h2 = v;
      else
        h3 = v;
      if (first == 4) first = slot;
      ++out;
      if (slot == 3) {
        int2* b = reinterpret_cast<int2*>(out - 4);
        if (first == 0) {
          store_sector(b, make_int4(h0.x, h0.y, h1.x, h1.y), make_int4(h2.x, h2.y, h3.x, h3.y));
        } else {  // the query's first sector: slots first .. 3
          if (first <= 1) b[1] = h1;
          if (first <= 2) b[2] = h2;
          b[3] = h3;
        }
        first = 4;
      }
    } else {
      out->index = i;
      out->distance = d;
      ++out;
    }
  }
  // the query's last, incomplete sector
  __device__ __forceinline__ void finish() {
    if constexpr (kStaged) {
      if (first == 4) return;
      const int end = slot_of(out);  // 1 .. 3: a complete sector has left already
      int2* b = reinterpret_cast<int2*>(out - end);
      if (first <= 0 && end > 0) b[0] = h0;
      if (first <= 1 && end > 1) b[1] = h1;
      if (first <= 2 && end > 2) b[2] = h2;
      first = 4;
    }
  }
};

}  // namespace pico

namespace pico {

// ---------------------------------------------------------------- exact nn over the search image (fat.cu)
// search_nn (search_visitor.hpp:41-65) that also notices when a second point attains the current
// best distance. The reference keeps the FIRST point visited at the minimum distance, and "first"
// depends on its visit order inside the collapsed subtrees, which a linear scan does not follow. A
// query whose final best is attained twice is therefore re-run by the order-exact traversal
// (traverse_packed); every other query has ONE point at the minimum distance, which any exact
// search returns.
template <typename T>
struct VisitNnTie {
  T best = Limits<T>::max();
  int idx = -1;
  bool tie = false;
  __device__ __forceinline__ void visit(int i, T d) {
    if (best > d) {
      best = d;
      idx = i;
      tie = false;
    } else if (best == d) {
      tie = true;
    }
  }
};

// ---------------------------------------------------------------- slot stack of the nn traversal
// Three words per slot, `[slot][thread]` in shared memory: the bank of an access is the thread's lane whatever
// its slot, so lanes that push and pop at different depths never conflict (per-thread local memory puts
// every lane's word in a line of its own: 158 M of the 344 M L1 sectors of the round-1 kernel were stack
// words, profiles/r1/kernels_v9_summary.txt). Slots beyond kSharedSlots spill to local memory (3 % of the
// cfg2 queries go deeper than four).
//   far child : tag = node | split_dim << 30, x = box distance, y = offset on split_dim inside it
//   restore   : tag = 3 << 30 | split_dim,    x = offset to put back when the subtree is done
// The reference sets node_box_offset_[split_dim] before it enters the far child and puts the old value back
// after (kd_tree_search.hpp:93-103); a far child that is taken leaves a restore record in its slot.
constexpr int kSharedSlots = 4;
constexpr uint32_t kTagRestore = 3u << 30, kTagNodeMask = (1u << 30) - 1u;

template <typename T, int THREADS>
struct SlotStack {
  uint32_t (*tag)[THREADS];  // shared [kSharedSlots][THREADS]
  T (*x)[THREADS];
  T (*y)[THREADS];
  uint32_t ltag[kLocalStack];  // spill
  T lx[kLocalStack], ly[kLocalStack];
  __device__ __forceinline__ void push(int sp, uint32_t t, T a, T b) {
    if (sp < kSharedSlots) {
      tag[sp][threadIdx.x] = t;
      x[sp][threadIdx.x] = a;
      y[sp][threadIdx.x] = b;
    } else {
      ltag[sp - kSharedSlots] = t;
      lx[sp - kSharedSlots] = a;
      ly[sp - kSharedSlots] = b;
    }
  }
  __device__ __forceinline__ void pop(int sp, uint32_t& t, T& a, T& b) const {
    if (sp < kSharedSlots) {
      t = tag[sp][threadIdx.x];
      a = x[sp][threadIdx.x];
      b = y[sp][threadIdx.x];
    } else {
      t = ltag[sp - kSharedSlots];
      a = lx[sp - kSharedSlots];
      b = ly[sp - kSharedSlots];
    }
  }
};

// Exact nearest neighbour, metric_l2_squared, sdim <= 3, for ONE query held in registers.
//
//   fat        search image: `nodes` with every subtree of <= fat_limit points collapsed to a leaf (fat.cu), or
//              the real `nodes` (FAT = false)
//   far_nodes  the array far children are traversed in (`fat` again, or the real `nodes`)
//
// 1. first descent, no frames (kd_tree_search.hpp:60-88 without the recursion), keeping the last NREC strict
//    prefix minima of the far-child offsets met on the way {value, branch node};
// 2. scan of the first leaf -> best. FAT: reach = best widened by 2^-13 of the first best, so that a far_dist
//    that rounding left a few ulps above the exact box distance (it is a running sum with one rounding per
//    level) can never hide a point AT the best distance — ties are always seen, and a query with a tie is
//    re-run in the reference's visit order (VisitNnTie). Not FAT: reach = best, the reference's own test, and
//    the visit order is the reference's (same argument as PRIME == kPrimeFirstLeaf of traverse_packed);
// 3. the far children of the first path that can matter are those with offset <= reach; the topmost of them is
//    necessarily a strict prefix minimum (everything above it is > reach >= it). If the newest record is
//    already > reach no far child matters and the query is done. Otherwise the second walk
//    (kd_tree_search.hpp:89-103 with an explicit stack) starts at the oldest recorded node whose value is
//    <= reach — from the root only when the evicted record would qualify as well. On the first path
//    node_box_distance and every offset are zero, so starting anywhere on it needs no other state.
template <typename T, int DIM, int NREC, bool FAT, typename Stack>
__device__ __forceinline__ void traverse_nn(const typename NodeOf<T>::type* __restrict__ fat,
                                            const typename NodeOf<T>::type* __restrict__ far_nodes,
                                            const typename Vec4Of<T>::type* __restrict__ pts4, const T (&q)[DIM],
                                            Stack& stack, VisitNnTie<T>& vis) {
  constexpr int metric = PICO_B200_METRIC_L2_SQUARED;
  constexpr int R = NREC > 0 ? NREC : 1;
  T a, b;
  uint32_t right, sd;
  int lb, le;
  uint32_t node = 0;
  T pm_val[R];
  uint32_t pm_node[R];
  T pm_evicted = Limits<T>::max();
#pragma unroll
  for (int r = 0; r < R; ++r) {
    pm_val[r] = Limits<T>::max();
    pm_node[r] = 0;
  }
  load_node(fat, node, a, b, right, sd, lb, le);
  while (sd != PICO_B200_LEAF) {
    T v = q[0];
#pragma unroll
    for (int j = 1; j < DIM; ++j)
      if (sd == (uint32_t)j) v = q[j];
    const bool go_left = sub_rn(sub_rn(add_rn(a, b), v), v) > T(0);
    if (NREC > 0) {
      const T t = sub_rn(go_left ? b : a, v);
      const T new_off = mul_rn(t, t);
      if (new_off < pm_val[0]) {
        pm_evicted = pm_val[R - 1];
#pragma unroll
        for (int r = R - 1; r > 0; --r) {
          pm_val[r] = pm_val[r - 1];
          pm_node[r] = pm_node[r - 1];
        }
        pm_val[0] = new_off;
        pm_node[0] = node;
      }
    }
    node = go_left ? node + 1 : right;
    load_node(fat, node, a, b, right, sd, lb, le);
  }
  const uint32_t primed_leaf = node;
  for (int i = lb; i < le; ++i) {
    const typename Vec4Of<T>::type p = ldg4(pts4 + i);
    T d = metric_first(metric, q[0], p.x);
    if (DIM > 1) d = metric_fold(metric, d, q[DIM > 1 ? 1 : 0], p.y, 1);
    if (DIM > 2) d = metric_fold(metric, d, q[DIM > 2 ? 2 : 0], p.z, 2);
    vis.visit(index_of(p), d);
  }
  const T margin = FAT ? mul_rn(vis.best, sizeof(T) == 4 ? T(1.2207031e-4) : T(2.2737368e-13)) : T(0);
  T reach = FAT ? add_rn(vis.best, margin) : vis.best;
  node = 0;
  if (NREC > 0) {
    if (reach < pm_val[0]) return;
    bool found = false;
#pragma unroll
    for (int r = R - 1; r >= 0; --r) {
      if (!found && pm_val[r] <= reach) {
        found = true;
        node = (r == R - 1 && pm_evicted <= reach) ? 0u : pm_node[r];
      }
    }
  }
  T off[DIM];
#pragma unroll
  for (int j = 0; j < DIM; ++j) off[j] = T(0);
  T node_dist = T(0);
  int sp = 0;
  const typename NodeOf<T>::type* cur = fat;
  for (;;) {
    load_node(cur, node, a, b, right, sd, lb, le);
    while (sd != PICO_B200_LEAF) {
      // (selects, not indexed stores: a run-time index would move off[] to local memory)
      T v = q[0], old = off[0];
#pragma unroll
      for (int j = 1; j < DIM; ++j) {
        v = (sd == (uint32_t)j) ? q[j] : v;
        old = (sd == (uint32_t)j) ? off[j] : old;
      }
      const bool go_left = sub_rn(sub_rn(add_rn(a, b), v), v) > T(0);
      const T t = sub_rn(go_left ? b : a, v);
      const T new_off = mul_rn(t, t);
      const uint32_t far = go_left ? right : node + 1;
      node = go_left ? node + 1 : right;
      const T far_dist = add_rn(sub_rn(node_dist, old), new_off);
      if (reach >= far_dist) {
        stack.push(sp, far | (sd << 30), far_dist, new_off);
        ++sp;
      }
      load_node(cur, node, a, b, right, sd, lb, le);
    }
    if (node == primed_leaf) le = lb;
    if (lb < le) {
      for (int i = lb; i < le; ++i) {
        const typename Vec4Of<T>::type p = ldg4(pts4 + i);
        T d = metric_first(metric, q[0], p.x);
        if (DIM > 1) d = metric_fold(metric, d, q[DIM > 1 ? 1 : 0], p.y, 1);
        if (DIM > 2) d = metric_fold(metric, d, q[DIM > 2 ? 2 : 0], p.z, 2);
        vis.visit(index_of(p), d);
      }
      reach = FAT ? add_rn(vis.best, margin) : vis.best;
    }
    bool found = false;
    while (sp > 0) {
      --sp;
      uint32_t tag;
      T x, y;
      stack.pop(sp, tag, x, y);
      const uint32_t hi = tag >> 30;
      if (hi == 3u) {  // the far subtree of this slot is done: put its offset back
#pragma unroll
        for (int j = 0; j < DIM; ++j) off[j] = ((tag & 3u) == (uint32_t)j) ? x : off[j];
        continue;
      }
      if (reach >= x) {
        node = tag & kTagNodeMask;
        node_dist = x;
        T old = off[0];
#pragma unroll
        for (int j = 1; j < DIM; ++j) old = (hi == (uint32_t)j) ? off[j] : old;
#pragma unroll
        for (int j = 0; j < DIM; ++j) off[j] = (hi == (uint32_t)j) ? y : off[j];
        if (sp > 0) {  // something is still pending below: it must see the offsets as they are now
          stack.push(sp, kTagRestore | hi, old, T(0));
          ++sp;
        }
        found = true;
        break;
      }
    }
    if (!found) return;
    cur = far_nodes;
  }
}

// ---------------------------------------------------------------- far subtrees as work items (nn_split_kernel)
// The order-exact traversal below one far child: the main loop of traverse_packed started at `node` with the box
// distance and the offsets that hold there. reach = best widened by `margin` (see VisitNnTie: ties must be seen).
template <typename T, int DIM, typename Stack>
__device__ __forceinline__ void traverse_subtree(const typename NodeOf<T>::type* __restrict__ nodes,
                                                 const typename Vec4Of<T>::type* __restrict__ pts4, const T (&q)[DIM],
                                                 uint32_t node, T node_dist, T (&off)[DIM], T margin, Stack& stack,
                                                 VisitNnTie<T>& vis) {
  constexpr int metric = PICO_B200_METRIC_L2_SQUARED;
  T reach = add_rn(vis.best, margin);
  int sp = 0;
  for (;;) {
    T a, b;
    uint32_t right, sd;
    int lb, le;
    load_node(nodes, node, a, b, right, sd, lb, le);
    while (sd != PICO_B200_LEAF) {
      T v = q[0], old = off[0];
#pragma unroll
      for (int j = 1; j < DIM; ++j) {
        v = (sd == (uint32_t)j) ? q[j] : v;
        old = (sd == (uint32_t)j) ? off[j] : old;
      }
      const bool go_left = sub_rn(sub_rn(add_rn(a, b), v), v) > T(0);
      const T t = sub_rn(go_left ? b : a, v);
      const T new_off = mul_rn(t, t);
      const uint32_t far = go_left ? right : node + 1;
      node = go_left ? node + 1 : right;
      const T far_dist = add_rn(sub_rn(node_dist, old), new_off);
      if (reach >= far_dist) {
        T snap[DIM];
#pragma unroll
        for (int j = 0; j < DIM; ++j) snap[j] = (sd == (uint32_t)j) ? new_off : off[j];
        stack.push(sp, far, far_dist, snap);
        ++sp;
      }
      load_node(nodes, node, a, b, right, sd, lb, le);
    }
    if (lb < le) {
      for (int i = lb; i < le; ++i) {
        const typename Vec4Of<T>::type p = ldg4(pts4 + i);
        T d = metric_first(metric, q[0], p.x);
        if (DIM > 1) d = metric_fold(metric, d, q[DIM > 1 ? 1 : 0], p.y, 1);
        if (DIM > 2) d = metric_fold(metric, d, q[DIM > 2 ? 2 : 0], p.z, 2);
        vis.visit(index_of(p), d);
      }
      reach = add_rn(vis.best, margin);
    }
    bool found = false;
    while (sp > 0) {
      --sp;
      T d;
      uint32_t n;
      T snap[DIM];
      stack.pop(sp, n, d, snap);
      if (reach >= d) {
        node = n;
        node_dist = d;
#pragma unroll
        for (int j = 0; j < DIM; ++j) off[j] = snap[j];
        found = true;
        break;
      }
    }
    if (!found) return;
  }
}

// ---------------------------------------------------------------- box search, thread-per-box, sdim <= 3
// search_box::operator() + report_node / report_left / report_right (kd_tree_search.hpp:270-372) for euclidean
// spaces, one thread per box: the running cell box (kd_tree_search.hpp:296-306 narrows one bound before it looks
// at a child and puts it back after) lives in registers, the recursion on a per-thread stack of
// {node, stage, saved bound}. Reports come in the reference's depth-first order: a cell inside the query is
// reported whole — its points are one contiguous run of the leaf-ordered arrays —, a leaf point by point
// (box.hpp:31-47, bounds inclusive). out == nullptr: count only.
struct BoxSlot {
  uint32_t node_stage;  // node | stage << 30
};

// The contiguous index range of a contained subtree (report_node) goes into the result row with 16-byte stores
// once the row position is aligned; single hits are written as they come. (Collecting single hits per 16 bytes in
// registers, as VisitRadiusFill does for its 8-byte records, cost the box kernel more in selects and registers than
// the stores it saved: 13.0 -> 16.0 ms per million boxes.)
__device__ __forceinline__ void copy_index_range(int32_t* __restrict__ dst, const int32_t* __restrict__ src, int n) {
  int i = 0;
  for (; i < n && (reinterpret_cast<uintptr_t>(dst + i) & 15) != 0; ++i) dst[i] = __ldg(src + i);
  for (; i + 4 <= n; i += 4)
    *reinterpret_cast<int4*>(dst + i) = make_int4(__ldg(src + i), __ldg(src + i + 1), __ldg(src + i + 2), __ldg(src + i + 3));
  for (; i < n; ++i) dst[i] = __ldg(src + i);
}

template <typename T, int DIM>
__device__ __forceinline__ uint32_t traverse_box_thread(const typename NodeOf<T>::type* __restrict__ nodes,
                                                        const typename Vec4Of<T>::type* __restrict__ pts4,
                                                        const int32_t* __restrict__ indices, const T (&qmin)[DIM],
                                                        const T (&qmax)[DIM], const T* __restrict__ root_box,
                                                        int32_t* __restrict__ out) {
  T bmin[DIM], bmax[DIM];
#pragma unroll
  for (int j = 0; j < DIM; ++j) {
    bmin[j] = root_box[j];
    bmax[j] = root_box[DIM + j];
  }
  uint32_t st_node[kLocalStack];
  T st_saved[kLocalStack];
  uint32_t count = 0;
  int sp = 1;
  st_node[0] = 0;
  while (sp > 0) {
    const uint32_t ns = st_node[sp - 1];
    const uint32_t node = ns & kTagNodeMask, stage = ns >> 30;
    T a, b;
    uint32_t right, sd;
    int lb, le;
    load_node(nodes, node, a, b, right, sd, lb, le);
    if (sd == PICO_B200_LEAF) {
      for (int i = lb; i < le; ++i) {
        const typename Vec4Of<T>::type p = ldg4(pts4 + i);
        bool in = !(qmin[0] > p.x || qmax[0] < p.x);
        if (DIM > 1) in = in && !(qmin[DIM > 1 ? 1 : 0] > p.y || qmax[DIM > 1 ? 1 : 0] < p.y);
        if (DIM > 2) in = in && !(qmin[DIM > 2 ? 2 : 0] > p.z || qmax[DIM > 2 ? 2 : 0] < p.z);
        if (in) {
          if (out) out[count] = index_of(p);
          ++count;
        }
      }
      --sp;
      continue;
    }
    if (stage == 2u) {  // both children done: the lower bound narrowed for the right child goes back
      const T saved = st_saved[sp - 1];
#pragma unroll
      for (int j = 0; j < DIM; ++j) bmin[j] = (sd == (uint32_t)j) ? saved : bmin[j];
      --sp;
      continue;
    }
    // stage 0: narrow max to left_max and look at the left child;
    // stage 1: put max back, narrow min to right_min and look at the right child
    const uint32_t child = stage == 0u ? node + 1 : right;
    T keep;
    if (stage == 0u) {
      keep = bmax[0];
#pragma unroll
      for (int j = 1; j < DIM; ++j) keep = (sd == (uint32_t)j) ? bmax[j] : keep;
#pragma unroll
      for (int j = 0; j < DIM; ++j) bmax[j] = (sd == (uint32_t)j) ? a : bmax[j];
    } else {
      const T saved = st_saved[sp - 1];
      keep = bmin[0];
#pragma unroll
      for (int j = 1; j < DIM; ++j) keep = (sd == (uint32_t)j) ? bmin[j] : keep;
#pragma unroll
      for (int j = 0; j < DIM; ++j) {
        bmax[j] = (sd == (uint32_t)j) ? saved : bmax[j];
        bmin[j] = (sd == (uint32_t)j) ? b : bmin[j];
      }
    }
    st_saved[sp - 1] = keep;
    st_node[sp - 1] = node | ((stage + 1u) << 30);
    // query.contains(box_) := contains(box.min) && contains(box.max), box.hpp:44-47
    bool contained = true;
#pragma unroll
    for (int j = 0; j < DIM; ++j)
      contained = contained && !(qmin[j] > bmin[j] || qmax[j] < bmin[j]) && !(qmin[j] > bmax[j] || qmax[j] < bmax[j]);
    if (contained) {
      // report_node (kd_tree_search.hpp:336-372): leftmost leaf's begin .. rightmost leaf's end
      uint32_t nl = child, nr = child, tr, tsd;
      int rb, re, d0, d1;
      T ta, tb;
      load_node(nodes, nl, ta, tb, tr, tsd, rb, d0);
      uint32_t rsd = tsd, rright = tr;
      re = d0;
      while (tsd != PICO_B200_LEAF) {
        ++nl;
        load_node(nodes, nl, ta, tb, tr, tsd, rb, d0);
      }
      while (rsd != PICO_B200_LEAF) {
        nr = rright;
        load_node(nodes, nr, ta, tb, rright, rsd, d1, re);
      }
      if (out) copy_index_range(out + count, indices + rb, re - rb);
      count += (uint32_t)(re - rb);
    } else {
      // intersects_left / intersects_right, kd_tree_search.hpp:310-328
      T lo = qmin[0], hi = qmax[0];
#pragma unroll
      for (int j = 1; j < DIM; ++j) {
        lo = (sd == (uint32_t)j) ? qmin[j] : lo;
        hi = (sd == (uint32_t)j) ? qmax[j] : hi;
      }
      const bool intersects = stage == 0u ? (lo <= a) : (hi >= b);
      if (intersects) {
        st_node[sp] = child;  // stage 0
        ++sp;
      }
    }
  }
  return count;
}

}  // namespace pico
