// common.cuh — shared declarations of libpico_b200.so (host + device).
//
// Data layout in HBM:
//   nodes   : Node<T>[n_nodes], pre-order, 16 B (f32) / 32 B (f64) records
//   pts4    : Vec4<T>[n] for sdim <= 3, LEAF ORDER, .w carries the original index
//   ptsN    : T[n * sdim] row-major for sdim > 3, LEAF ORDER
//   indices : int32[n], leaf position -> original index (kd_tree_data::indices)
//   outer   : T[n_nodes][2] = {left_min, right_max}, topological metrics only
//   spans   : uint2[n_nodes] = {first point, number of points} below each node, sdim > 3 only
//             (leaf order makes every subtree one contiguous run; feeds the subtree distance cache)
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>

#include "../../include/pico_b200.h"

namespace pico {

// ---------------------------------------------------------------- errors
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);

#define PICO_CUDA(expr)                                                                         \
  do {                                                                                          \
    cudaError_t e__ = (expr);                                                                   \
    if (e__ != cudaSuccess) {                                                                   \
      return ::pico::fail(                                                                      \
          e__ == cudaErrorMemoryAllocation ? PICO_B200_ERR_OUT_OF_MEMORY : PICO_B200_ERR_CUDA,  \
          std::string(#expr) + ": " + cudaGetErrorString(e__));                                 \
    }                                                                                           \
  } while (0)

#define PICO_TRY(expr)           \
  do {                           \
    int rc__ = (expr);           \
    if (rc__ != 0) return rc__;  \
  } while (0)

// ---------------------------------------------------------------- node / point records
template <typename T>
struct NodeOf;
template <>
struct NodeOf<float> {
  using type = pico_b200_node_f32;
};
template <>
struct NodeOf<double> {
  using type = pico_b200_node_f64;
};

template <typename T>
struct Vec4Of;
template <>
struct Vec4Of<float> {
  using type = float4;
};
template <>
struct Vec4Of<double> {
  using type = double4;
};

template <typename T>
struct Neighbor {
  int32_t index;
  T distance;
};
static_assert(sizeof(Neighbor<float>) == 8, "neighbor<int,float> is 8 bytes");
static_assert(sizeof(Neighbor<double>) == 16, "neighbor<int,double> is 16 bytes");

template <typename T>
struct Limits;
template <>
struct Limits<float> {
  __host__ __device__ static constexpr float max() { return 3.402823466e+38f; }
};
template <>
struct Limits<double> {
  __host__ __device__ static constexpr double max() { return 1.7976931348623158e+308; }
};

constexpr int kMaxPackedDim = 3;   // sdim <= 3 uses the Vec4 layout
constexpr int kLocalStack = 64;    // traversal stack kept in per-thread local memory
constexpr int kSmBlocks = 148;     // B200 SM count (grid sizing; queried at runtime too)
constexpr int kFatLeafPoints = 0;  // subtrees of at most this many points are one leaf of the search image (fat.cu); 0 = none

}  // namespace pico

// ---------------------------------------------------------------- the handle
// What the last batches looked like to the query ordering (search.cu: tile_order_kernel). Ordering never changes a
// result, so a stale or racy hint costs time at worst.
struct pico_b200_order_hint {
  std::atomic<int> state{0};               // 0 unknown, 1 batches arrive locally coherent, 2 they need the global sort
  std::atomic<unsigned> calls{0};
  unsigned long long* h_stat = nullptr;    // pinned: tiles << 40 | distinct coarse cells summed over the tiles, last measured batch
  unsigned long long* d_stat = nullptr;
};

struct pico_b200_tree {
  int device = 0;
  int scalar = PICO_B200_F32;
  int metric = PICO_B200_METRIC_L2_SQUARED;
  size_t n = 0, sdim = 0, n_nodes = 0, n_leaves = 0, height = 0, max_leaf_points = 0;
  void* d_nodes = nullptr;
  void* d_pts = nullptr;       // pts4 or ptsN (see above)
  int32_t* d_indices = nullptr;
  void* d_root_box = nullptr;  // min[sdim] then max[sdim], device copy
  void* d_outer = nullptr;     // topological metrics: {left_min, right_max}[n_nodes] (kd_tree_node.hpp:52-59)
  uint2* d_spans = nullptr;    // row storage (sdim > 3): {first point, point count} below every node
  void* d_fat_nodes = nullptr; // packed trees: `nodes` with small subtrees collapsed into leaves (fat.cu), or null
  int fat_limit = 0;           // points per collapsed subtree (0 = no search image)
  mutable pico_b200_order_hint order_hint;
  double root_box_host[2 * 4] = {0};  // first min(sdim,4) dims, as double, for query ordering
  double build_ms = 0.0;
  size_t device_bytes = 0;
  int sm_count = pico::kSmBlocks;
  size_t scalar_size() const { return scalar == PICO_B200_F64 ? 8 : 4; }
  size_t node_size() const { return scalar == PICO_B200_F64 ? 32 : 16; }
  bool packed() const { return sdim <= (size_t)pico::kMaxPackedDim; }
  bool keep_outer = false;     // trees of a kd_forest: euclidean metric, but the priority search needs all four bounds
  bool topological() const { return metric >= PICO_B200_METRIC_SO2 && metric <= PICO_B200_METRIC_CUSTOM_TOPOLOGICAL; }
  bool custom_metric() const { return metric >= PICO_B200_METRIC_CUSTOM_TOPOLOGICAL; }
  size_t outer_bytes() const { return (topological() || keep_outer) ? n_nodes * 2 * scalar_size() : 0; }
  size_t spans_bytes() const { return packed() ? 0 : n_nodes * sizeof(uint2); }
  size_t pts_bytes() const { return packed() ? n * 4 * scalar_size() : n * sdim * scalar_size(); }
};

namespace pico {

// build.cu
template <typename T>
int build_tree(pico_b200_tree* t, const T* h_pts, size_t stride, int rule, int stop_kind, size_t stop_value,
               const T* bounds_min, const T* bounds_max);
template <typename T>
int upload_tree(pico_b200_tree* t, const T* h_pts, size_t stride, const void* nodes, size_t n_nodes,
                const int32_t* indices, const T* root_box, const T* outer_bounds);

// fat.cu: (re)builds t->d_fat_nodes from t->d_nodes; every way of making a tree ends with it
int build_fat_nodes(pico_b200_tree* t, cudaStream_t st);
// search.cu: frees a host result of the ragged searches (one big block is kept for the next call)
void release_result(void* p);

// search.cu
int order_state(const pico_b200_tree* t);  // pico_b200_order_hint::state after looking at the last measured batch
int set_thread_stream(void* stream, bool has);
int profile_begin();
int profile_end(double* ms, uint64_t* launches);
template <typename T>
int knn_batch(const pico_b200_tree* t, const T* q, size_t nq, size_t stride, size_t k, double e, Neighbor<T>* out,
              unsigned flags, pico_b200_search_stats* stats);
template <typename T>
int leaf_scan_profile(const pico_b200_tree* t, const T* d_q, size_t nq, size_t stride, Neighbor<T>* d_out, int repeats,
                      double* descend_ms, double* scan_ms, uint64_t* scan_bytes);
template <typename T>
int radius_batch(const pico_b200_tree* t, const T* q, size_t nq, size_t stride, double radius, double e,
                 uint64_t* offsets_out, void** out, unsigned flags, pico_b200_search_stats* stats);
template <typename T>
int box_batch(const pico_b200_tree* t, const T* mins, const T* maxs, size_t nb, size_t stride,
              uint64_t* offsets_out, int32_t** out, unsigned flags, pico_b200_search_stats* stats);

}  // namespace pico
