// forest.cuh — best-bin-first traversal of ONE tree of a kd_forest over the flat pre-order node array.
//
// Device-side counterpart of priority_search_nearest_euclidean
// (examples/pico_understory/pico_understory/internal/kd_tree_priority_search.hpp:24-142) as
// kd_forest::search_nearest drives it, one tree after the other with a shared visitor
// (examples/pico_understory/pico_understory/kd_forest.hpp:91-120). SURVEY.md §8 f4.
//
// The arithmetic of the search (reflection, branch step, queue order) is written `__host__ __device__`: the very
// same source is checked on the CPU against the reference fixtures (tests/cpp/forest_host.cpp ->
// tests/test_forest_core.py, scalar driver `priority_search_tree` below) and runs inside the warp-per-query
// kernel of forest.cu (handle + C-ABI pico_b200_forest_*).
//
// Reference recursion -> iteration. One "descent" of the reference (:66-125) walks from a queued node to a
// leaf through the nearer children, scans the leaf, and while the recursion unwinds queues every farther
// child whose box distance is still below visitor.max(). The box distance handed to the nearer child is the
// parent's (:114), so every farther child of one descent is measured against the distance of the node the
// descent started from; and all their `max() > distance` tests (:122) happen after the single leaf scan of
// that descent, against one and the same max(). Hence: record {far child, distance} on the way down, scan,
// then queue the recorded ones that pass — same set, and the queue is ordered by a total order, so the
// insertion order does not matter.
#pragma once

#include <stddef.h>
#include <stdint.h>

#include "../../include/pico_b200.h"

#if defined(__CUDACC__)
#define PICO_FOREST_HD __host__ __device__ __forceinline__
#else
#define PICO_FOREST_HD inline
#endif

namespace pico {
namespace forest {

// Separate multiply / add / subtract, never contracted: intrinsics on the device, plain operators on the host
// (host translation units that include this header are compiled with -ffp-contract=off).
#if defined(__CUDA_ARCH__)
PICO_FOREST_HD float f_add(float a, float b) { return __fadd_rn(a, b); }
PICO_FOREST_HD float f_sub(float a, float b) { return __fsub_rn(a, b); }
PICO_FOREST_HD float f_mul(float a, float b) { return __fmul_rn(a, b); }
PICO_FOREST_HD double f_add(double a, double b) { return __dadd_rn(a, b); }
PICO_FOREST_HD double f_sub(double a, double b) { return __dsub_rn(a, b); }
PICO_FOREST_HD double f_mul(double a, double b) { return __dmul_rn(a, b); }
#else
PICO_FOREST_HD float f_add(float a, float b) { return a + b; }
PICO_FOREST_HD float f_sub(float a, float b) { return a - b; }
PICO_FOREST_HD float f_mul(float a, float b) { return a * b; }
PICO_FOREST_HD double f_add(double a, double b) { return a + b; }
PICO_FOREST_HD double f_sub(double a, double b) { return a - b; }
PICO_FOREST_HD double f_mul(double a, double b) { return a * b; }
#endif

template <typename T>
struct NodeOfT;
template <>
struct NodeOfT<float> {
  using type = pico_b200_node_f32;
};
template <>
struct NodeOfT<double> {
  using type = pico_b200_node_f64;
};

// One tree of the forest: flat pre-order nodes (left child = i + 1), the two extra bounds of
// kd_tree_node_topological per node ({left_min, right_max}, kd_tree_node.hpp:52-59), the index permutation and
// the Householder-reflected copy of the point set in ORIGINAL order (rkd_tree_hh_data.hpp:51-63), row-major.
template <typename T>
struct TreeView {
  const typename NodeOfT<T>::type* nodes;
  const T* outer;
  const int32_t* indices;
  const T* points;
  size_t stride;
  uint32_t sdim;
};

// rkd_tree_hh_data::rotate_point, rkd_tree_hh_data.hpp:79-90: y = x - 2 (r . x) r, dot accumulated in index order.
template <typename T>
PICO_FOREST_HD void householder(const T* r, uint32_t sdim, const T* x, T* y) {
  T dot = T(0);
  for (uint32_t i = 0; i < sdim; ++i) dot = f_add(dot, f_mul(r[i], x[i]));
  dot = f_mul(dot, T(2));
  for (uint32_t i = 0; i < sdim; ++i) y[i] = f_sub(x[i], f_mul(dot, r[i]));
}

template <typename T>
struct Entry {
  T dist;
  uint32_t node;
};

// std::priority_queue<pair<scalar, node const*>, ..., std::greater<>> (kd_tree_priority_search.hpp:136-140) as a
// binary heap in caller-provided storage. Equal distances: the reference compares node addresses; here the
// smaller pre-order number wins (the oracle's rule, oracle/pico_oracle_impl.inc po_qless).
template <typename T>
struct MinQueue {
  Entry<T>* a;
  uint32_t n, cap;
  bool overflow;
  PICO_FOREST_HD static bool less(const Entry<T>& x, const Entry<T>& y) {
    return x.dist < y.dist || (!(y.dist < x.dist) && x.node < y.node);
  }
  PICO_FOREST_HD void clear() { n = 0; }
  PICO_FOREST_HD bool empty() const { return n == 0; }
  PICO_FOREST_HD const Entry<T>& top() const { return a[0]; }
  PICO_FOREST_HD void push(T dist, uint32_t node) {
    if (n == cap) {  // never silently drop a candidate: the caller sizes the storage or reports the overflow
      overflow = true;
      return;
    }
    Entry<T> e;
    e.dist = dist;
    e.node = node;
    uint32_t i = n++;
    while (i > 0 && less(e, a[(i - 1) / 2])) {
      a[i] = a[(i - 1) / 2];
      i = (i - 1) / 2;
    }
    a[i] = e;
  }
  PICO_FOREST_HD void pop() {
    const Entry<T> e = a[--n];
    uint32_t i = 0;
    for (;;) {
      uint32_t c = 2 * i + 1;
      if (c >= n) break;
      if (c + 1 < n && less(a[c + 1], a[c])) ++c;
      if (!less(a[c], e)) break;
      a[i] = a[c];
      i = c;
    }
    if (n) a[i] = e;
  }
};

template <typename T>
PICO_FOREST_HD void load(const pico_b200_node_f32* nodes, uint32_t i, T& a, T& b, uint32_t& right, uint32_t& sd,
                         int32_t& lb, int32_t& le) {
  const pico_b200_node_f32 nd = nodes[i];
  a = nd.a.left_max;
  b = nd.b.right_min;
  lb = nd.a.begin_idx;
  le = nd.b.end_idx;
  right = nd.right;
  sd = nd.split_dim;
}
template <typename T>
PICO_FOREST_HD void load(const pico_b200_node_f64* nodes, uint32_t i, T& a, T& b, uint32_t& right, uint32_t& sd,
                         int32_t& lb, int32_t& le) {
  const pico_b200_node_f64 nd = nodes[i];
  a = nd.a.left_max;
  b = nd.b.right_min;
  lb = (int32_t)nd.a.begin_idx;
  le = (int32_t)nd.b.end_idx;
  right = nd.right;
  sd = nd.split_dim;
}

// One branch of a descent (kd_tree_priority_search.hpp:88-119): the nearer child is entered, the farther one is
// recorded with the box distance  parent_dist - old_offset + new_offset  (one rounding each, left to right). The
// nearer child inherits parent_dist (:114). Shared by the scalar core below and the warp kernel (forest.cu).
//   a = left_max, b = right_min, v = query coordinate on the split dimension
template <typename T>
PICO_FOREST_HD void branch_step(T a, T b, T left_min, T right_max, T v, T parent_dist, uint32_t node, uint32_t right,
                                uint32_t& first, uint32_t& second, T& second_dist) {
  T old_offset, new_offset;
  if (f_sub(f_sub(f_add(a, b), v), v) > T(0)) {  // :92-101
    first = node + 1;
    second = right;
    const T t0 = f_sub(left_min, v);
    old_offset = (v > left_min) ? T(0) : f_mul(t0, t0);
    const T t1 = f_sub(b, v);
    new_offset = f_mul(t1, t1);
  } else {  // :102-111
    first = right;
    second = node + 1;
    const T t0 = f_sub(right_max, v);
    old_offset = (v < right_max) ? T(0) : f_mul(t0, t0);
    const T t1 = f_sub(a, v);
    new_offset = f_mul(t1, t1);
  }
  second_dist = f_add(f_sub(parent_dist, old_offset), new_offset);  // :119
}

// priority_search_nearest_euclidean::operator() (:47-63) for one tree, metric_l2_squared.
//   q          the query already reflected into this tree's space (kd_forest.hpp:103)
//   vis        visitor with max() and visit(index, distance), shared by all trees of the forest
//   queue      storage for the priority queue; path: scratch for one descent, at least height + 1 entries
// Returns the number of leaves visited.
template <typename T, typename Visitor>
PICO_FOREST_HD uint32_t priority_search_tree(const TreeView<T>& tree, const T* q, size_t max_leaves_visited,
                                             Visitor& vis, MinQueue<T>& queue, Entry<T>* path) {
  uint32_t leaves_visited = 0;
  queue.clear();
  queue.push(T(0), 0u);
  while (!queue.empty()) {
    const Entry<T> top = queue.top();
    if (leaves_visited >= max_leaves_visited || vis.max() < top.dist) break;  // :53-56
    queue.pop();
    // ---- one descent (:66-125)
    uint32_t node = top.node, n_path = 0;
    T a, b;
    uint32_t right, sd;
    int32_t lb, le;
    load<T>(tree.nodes, node, a, b, right, sd, lb, le);
    while (sd != PICO_B200_LEAF) {
      uint32_t first, second;
      T dist;
      branch_step<T>(a, b, tree.outer[2 * (size_t)node], tree.outer[2 * (size_t)node + 1], q[sd], top.dist, node, right,
                     first, second, dist);
      path[n_path].node = second;
      path[n_path].dist = dist;
      ++n_path;
      node = first;
      load<T>(tree.nodes, node, a, b, right, sd, lb, le);
    }
    for (int32_t i = lb; i < le; ++i) {  // :68-73
      const int32_t idx = tree.indices[i];
      const T* p = tree.points + (size_t)idx * tree.stride;
      T d = T(0);
      for (uint32_t j = 0; j < tree.sdim; ++j) {  // metric.hpp:36-51,103-117
        const T t = f_sub(q[j], p[j]);
        d = f_add(d, f_mul(t, t));
      }
      vis.visit(idx, d);
    }
    for (uint32_t i = n_path; i-- > 0;)  // :122-124, deepest first like the unwinding recursion
      if (vis.max() > path[i].dist) queue.push(path[i].dist, path[i].node);
    ++leaves_visited;
  }
  return leaves_visited;
}

}  // namespace forest
}  // namespace pico
