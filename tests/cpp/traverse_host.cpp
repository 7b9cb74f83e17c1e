// traverse_host.cpp — TEST ONLY: compiles the thread-per-query traversals of pico_tree_b200/csrc/traverse.cuh
// (the product's device code) for the HOST, one "thread" at a time, so that their control flow — the
// search-image nn traversal with its prefix-minimum restart and tie detection, and the order-exact
// traverse_packed — can be checked against the oracle without a GPU (tests/test_traverse_host.py).
// The CUDA intrinsics the header uses are given their IEEE meaning here; the file is compiled with
// -ffp-contract=off like every other host translation unit that has to match the reference bit for bit.
// Built by tests/cpp/Makefile into tests/_bin/libtraverse_host.so. Not part of the product.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstring>

static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline long long __double_as_longlong(double d) { long long i; std::memcpy(&i, &d, 8); return i; }
static inline double __longlong_as_double(long long i) { double d; std::memcpy(&d, &i, 8); return d; }
// loads by record type (uint4 = f32 node, float4 = f32 point ...): statistics for tests/test_traverse_host.py
static thread_local unsigned long long g_loads16 = 0, g_loads_pts = 0;
static const uint3 threadIdx = {0, 0, 0};  // one "thread" at a time
static inline unsigned __activemask() { return 1u; }
template <typename V>
static inline V __shfl_sync(unsigned, V v, int) { return v; }
template <typename V>
static inline V __ldg(const V* p) {
  ++g_loads16;
  return *p;
}
static inline float4 __ldg(const float4* p) {
  ++g_loads_pts;
  return *p;
}

#include "../../pico_tree_b200/csrc/traverse.cuh"

// stack statistics (pushes, pops, deepest stack) for tests/test_traverse_host.py
static thread_local unsigned long long g_knn_visits = 0, g_knn_inserts = 0;
static thread_local unsigned long long g_push = 0, g_pop = 0, g_max_sp_sum = 0, g_sp_hist[8] = {0};
template <typename T>
struct CountingStack {
  uint32_t tag[pico::kSharedSlots][1];
  T x[pico::kSharedSlots][1], y[pico::kSharedSlots][1];
  pico::SlotStack<T, 1> s;
  int max_sp = 0;
  CountingStack() {
    s.tag = tag;
    s.x = x;
    s.y = y;
  }
  void push(int sp, uint32_t t, T a, T b) {
    if ((t >> 30) != 3u) ++g_push;
    if (sp + 1 > max_sp) max_sp = sp + 1;
    s.push(sp, t, a, b);
  }
  void pop(int sp, uint32_t& t, T& a, T& b) const {
    s.pop(sp, t, a, b);
    if ((t >> 30) != 3u) ++g_pop;
  }
};

namespace {

template <typename T, int DIM>
void nn_fat(const void* fat, const void* far_nodes, const void* pts4, const T* q, size_t nq, int nrec, int32_t* idx,
            T* dist, uint8_t* tie) {
  using namespace pico;
  using NodeT = typename NodeOf<T>::type;
  using V4 = typename Vec4Of<T>::type;
  for (size_t i = 0; i < nq; ++i) {
    T qq[DIM];
    for (int j = 0; j < DIM; ++j) qq[j] = q[i * DIM + j];
    VisitNnTie<T> vis;
    CountingStack<T> st;
    const NodeT* f = static_cast<const NodeT*>(fat);
    const NodeT* fr = static_cast<const NodeT*>(far_nodes);
    const V4* p4 = static_cast<const V4*>(pts4);
    if (f == fr && nrec >= 100) {  // no search image: the reference's own prune test and visit order
      if (nrec == 100)
        traverse_nn<T, DIM, 0, false>(f, fr, p4, qq, st, vis);
      else
        traverse_nn<T, DIM, 3, false>(f, fr, p4, qq, st, vis);
    } else if (nrec == 0) {
      traverse_nn<T, DIM, 0, true>(f, fr, p4, qq, st, vis);
    } else if (nrec == 1) {
      traverse_nn<T, DIM, 1, true>(f, fr, p4, qq, st, vis);
    } else {
      traverse_nn<T, DIM, 3, true>(f, fr, p4, qq, st, vis);
    }
    idx[i] = vis.idx;
    dist[i] = vis.best;
    tie[i] = vis.tie ? 1 : 0;
    g_max_sp_sum += st.max_sp;
    g_sp_hist[st.max_sp < 7 ? st.max_sp : 7]++;
  }
}

// order-exact search_nn / search_knn (k <= 16) through traverse_packed, PRIME as the kernels use it
template <typename T, int DIM>
void knn_packed(const void* nodes, const void* pts4, const T* q, size_t nq, int k, int n_points, int32_t* idx, T* dist) {
  using namespace pico;
  using NodeT = typename NodeOf<T>::type;
  using V4 = typename Vec4Of<T>::type;
  for (size_t i = 0; i < nq; ++i) {
    T qq[DIM];
    for (int j = 0; j < DIM; ++j) qq[j] = q[i * DIM + j];
    LocalStack<T, DIM, kLocalStack> st;
    if (k == 1) {
      VisitNn<T> vis;
      traverse_packed<T, DIM, true, kPrimeFirstLeaf>(static_cast<const NodeT*>(nodes), static_cast<const V4*>(pts4),
                                                     nullptr, qq, 0, false, T(1), st, vis);
      idx[i] = vis.idx;
      dist[i] = vis.best;
    } else {
      struct CountingKnn : VisitKnn<T, 16> {
        void visit(int i, T x) {
          ++g_knn_visits;
          if (this->d[15] > x) ++g_knn_inserts;
          VisitKnn<T, 16>::visit(i, x);
        }
      } vis;
      vis.init(k);
      traverse_packed<T, DIM, true, kPrimeBound>(static_cast<const NodeT*>(nodes), static_cast<const V4*>(pts4), nullptr,
                                                 qq, 0, false, T(1), st, vis, n_points, k);
      for (int j = 0; j < k; ++j) {
        idx[i * k + j] = vis.id[16 - k + j];
        dist[i * k + j] = vis.d[16 - k + j];
      }
    }
  }
}

}  // namespace

// thread-per-box traversal: count pass, then fill pass, per box
template <int DIM>
static unsigned long long box_f32(const void* nodes, const void* pts4, const int32_t* indices, const float* root_box,
                                  const float* mins, const float* maxs, size_t nb, uint64_t* offsets, int32_t* out) {
  using namespace pico;
  unsigned long long total = 0;
  for (size_t i = 0; i < nb; ++i) {
    float lo[DIM], hi[DIM];
    for (int j = 0; j < DIM; ++j) {
      lo[j] = mins[i * DIM + j];
      hi[j] = maxs[i * DIM + j];
    }
    offsets[i] = total;
    const uint32_t c = traverse_box_thread<float, DIM>(static_cast<const pico_b200_node_f32*>(nodes),
                                                       static_cast<const float4*>(pts4), indices, lo, hi, root_box,
                                                       out ? out + total : nullptr);
    total += c;
  }
  offsets[nb] = total;
  return total;
}

extern "C" {
// search_box through traverse_box_thread; out == nullptr: offsets only. Returns the number of reported indices.
unsigned long long host_box_f32(const void* nodes, const void* pts4, const int32_t* indices, const float* root_box,
                                const float* mins, const float* maxs, size_t nb, int sdim, uint64_t* offsets,
                                int32_t* out) {
  if (sdim == 2) return box_f32<2>(nodes, pts4, indices, root_box, mins, maxs, nb, offsets, out);
  return box_f32<3>(nodes, pts4, indices, root_box, mins, maxs, nb, offsets, out);
}
// node loads / f32 point loads since the last call
unsigned long long host_take_loads16() {
  const unsigned long long v = g_loads16;
  g_loads16 = 0;
  return v;
}
// {pushes, pops, sum of deepest stack, histogram of deepest stack 0..7+} since the last call
void host_take_stack_stats(unsigned long long* out11) {
  out11[0] = g_push; out11[1] = g_pop; out11[2] = g_max_sp_sum;
  for (int i = 0; i < 8; ++i) { out11[3 + i] = g_sp_hist[i]; g_sp_hist[i] = 0; }
  g_push = g_pop = g_max_sp_sum = 0;
}
void host_take_knn_stats(unsigned long long* out2) {
  out2[0] = g_knn_visits;
  out2[1] = g_knn_inserts;
  g_knn_visits = g_knn_inserts = 0;
}
unsigned long long host_take_point_loads() {
  const unsigned long long v = g_loads_pts;
  g_loads_pts = 0;
  return v;
}
void host_nn_fat_f32(const void* fat, const void* far_nodes, const void* pts4, const float* q, size_t nq, int sdim,
                     int nrec, int32_t* idx, float* dist, uint8_t* tie) {
  if (sdim == 2)
    nn_fat<float, 2>(fat, far_nodes, pts4, q, nq, nrec, idx, dist, tie);
  else
    nn_fat<float, 3>(fat, far_nodes, pts4, q, nq, nrec, idx, dist, tie);
}
void host_nn_fat_f64(const void* fat, const void* far_nodes, const void* pts4, const double* q, size_t nq, int sdim,
                     int nrec, int32_t* idx, double* dist, uint8_t* tie) {
  if (sdim == 2)
    nn_fat<double, 2>(fat, far_nodes, pts4, q, nq, nrec, idx, dist, tie);
  else
    nn_fat<double, 3>(fat, far_nodes, pts4, q, nq, nrec, idx, dist, tie);
}
void host_knn_packed_f32(const void* nodes, const void* pts4, const float* q, size_t nq, int sdim, int k, int n_points,
                         int32_t* idx, float* dist) {
  if (sdim == 2)
    knn_packed<float, 2>(nodes, pts4, q, nq, k, n_points, idx, dist);
  else
    knn_packed<float, 3>(nodes, pts4, q, nq, k, n_points, idx, dist);
}
}
