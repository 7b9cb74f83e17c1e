// traverse_warp.cuh — warp-per-query traversal for any spatial dimension / any k.
//
// All 32 lanes walk the tree together (node loads are warp-uniform broadcasts); a leaf is
// scanned with one lane per point, each lane accumulating its distance over the dimensions
// in ascending order (so the rounding equals the reference's sequential sum,
// metric.hpp:36-51), and candidates are handed to the visitor in leaf order so that
// "first visited wins" (search_visitor.hpp:55,107,141) is preserved.
//
// The reference's set/restore of node_box_offset_ (kd_tree_search.hpp:93-103) is kept
// literally: frames on a per-warp stack in global memory carry {far child, parent box
// distance, new offset, saved offset, split_dim | state}; the offset vector itself lives
// in shared memory next to the query.
#pragma once

#include "traverse.cuh"

namespace pico {

constexpr int kStackWindow = 16;  // frames mirrored in shared memory (power of two)

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <typename T>
struct WarpFrame {
  uint32_t far;
  uint32_t sd_state;  // split_dim | (ACTIVE << 31)
  T dist;             // box distance of the parent (PENDING) — unused once ACTIVE
  T new_off;
  T old_off;
};

// Point access: PACKED = Vec4 records with the index in .w (sdim <= 3); else rows + indices.
template <typename T, bool PACKED>
struct PointSet {
  const typename Vec4Of<T>::type* pts4;
  const T* rows;
  const int32_t* indices;
  int sdim;
  __device__ __forceinline__ T distance(int i, const T* q, int metric, int& index) const {
    T d = metric_init<T>(metric);
    if (PACKED) {
      const typename Vec4Of<T>::type p = ldg4(pts4 + i);
      d = metric_fold(metric, d, q[0], p.x, 0);
      if (sdim > 1) d = metric_fold(metric, d, q[1], p.y, 1);
      if (sdim > 2) d = metric_fold(metric, d, q[2], p.z, 2);
      index = index_of(p);
    } else {
      const T* p = rows + (size_t)i * sdim;
      for (int j = 0; j < sdim; ++j) d = metric_fold(metric, d, q[j], __ldg(p + j), j);
      index = __ldg(indices + i);
    }
    return d;
  }
  // is coordinate x of dimension j inside the query interval: box_base::contains(point), box.hpp:31-40
  // (inclusive); topological spaces: metric_box_map::contains, box.hpp:305-312 — an interval on the
  // circle with min > max wraps around (segment_s1::contains, segment.hpp:63-69)
  __device__ __forceinline__ static bool coord_inside(int metric, int j, T mn, T mx, T x) {
    if (!is_topological(metric)) return !(mn > x || mx < x);
    if (!dim_is_s1(metric, (uint32_t)j) || mn <= mx) return mn <= x && x <= mx;
    return x >= mn || x <= mx;
  }
  __device__ __forceinline__ bool inside(int i, const T* qmin, const T* qmax, int metric, int& index) const {
    bool ok = true;
    if (PACKED) {
      const typename Vec4Of<T>::type p = ldg4(pts4 + i);
      const T c[3] = {p.x, p.y, p.z};
      for (int j = 0; j < sdim; ++j) ok = ok && coord_inside(metric, j, qmin[j], qmax[j], c[j]);
      index = index_of(p);
    } else {
      const T* p = rows + (size_t)i * sdim;
      for (int j = 0; j < sdim; ++j) ok = ok && coord_inside(metric, j, qmin[j], qmax[j], __ldg(p + j));
      index = __ldg(indices + i);
    }
    return ok;
  }
};

// ---------------------------------------------------------------- leaf scan for row storage (sdim > 3)
template <typename T>
struct Vec16;
template <>
struct Vec16<float> {
  using type = float4;
  static constexpr int n = 4;
};
template <>
struct Vec16<double> {
  using type = double2;
  static constexpr int n = 2;
};

// The distance of a row is a fold over per-coordinate TERMS: term_j = (q_j - p_j)^2 or |q_j - p_j|,
// folded with +, max or min in dimension order (metric.hpp:36-51,131-145,158-174). The terms are
// independent of each other, the fold is a serial chain. Leaves are small, so a lane-per-point
// loop doing both leaves most lanes idle while the warp still pays every instruction: measured
// 364 M warp instructions per 128-D query, 70 % of them in that loop, issue slots 79 % busy.
// Split instead: ALL lanes compute terms (coalesced 16-byte loads straight from global memory, one
// chunk of a row per lane) into a shared-memory tile; then lane p folds the terms of row p in
// dimension order. Same operations on the same operands in the same order — bit-identical.
template <int METRIC>
__device__ __forceinline__ float term1(float q, float p) {
  const float t = sub_rn(q, p);
  return METRIC == PICO_B200_METRIC_L2_SQUARED ? mul_rn(t, t) : fabsf(t);
}
template <int METRIC>
__device__ __forceinline__ double term1(double q, double p) {
  const double t = sub_rn(q, p);
  return METRIC == PICO_B200_METRIC_L2_SQUARED ? mul_rn(t, t) : fabs(t);
}
template <int METRIC, typename T>
__device__ __forceinline__ T acc1(T d, T a) {
  if (METRIC == PICO_B200_METRIC_L2_SQUARED || METRIC == PICO_B200_METRIC_L1) return add_rn(d, a);
  if (METRIC == PICO_B200_METRIC_LPINF) return d < a ? a : d;
  return a < d ? a : d;
}
template <int METRIC>
__device__ __forceinline__ float4 term_vec(const float4& q, const float4& p) {
  return make_float4(term1<METRIC>(q.x, p.x), term1<METRIC>(q.y, p.y), term1<METRIC>(q.z, p.z),
                     term1<METRIC>(q.w, p.w));
}
template <int METRIC>
__device__ __forceinline__ double2 term_vec(const double2& q, const double2& p) {
  return make_double2(term1<METRIC>(q.x, p.x), term1<METRIC>(q.y, p.y));
}
template <int METRIC>
__device__ __forceinline__ float acc_vec(float d, const float4& s) {
  d = acc1<METRIC>(d, s.x);
  d = acc1<METRIC>(d, s.y);
  d = acc1<METRIC>(d, s.z);
  return acc1<METRIC>(d, s.w);
}
template <int METRIC>
__device__ __forceinline__ double acc_vec(double d, const double2& s) {
  d = acc1<METRIC>(d, s.x);
  return acc1<METRIC>(d, s.y);
}
__device__ __forceinline__ float4 ldg16(const float4* p) { return __ldg(p); }
__device__ __forceinline__ double2 ldg16(const double2* p) { return __ldg(p); }

// rows [base, base + rows_here) -> terms in the tile -> lane p < rows_here returns the fold of row p
template <int METRIC, typename T>
__device__ __forceinline__ T terms_then_fold(const T* __restrict__ rows, int sdim, int base, int rows_here,
                                             const T* sq, T* tile) {
  using VT = typename Vec16<T>::type;
  constexpr int V = Vec16<T>::n;
  const int lane = threadIdx.x & 31;
  const int sdimv = sdim / V;
  const int pitch = sdimv + 1;  // one 16-byte pad per row: rows start in different banks
  VT* tile_v = reinterpret_cast<VT*>(tile);
  const VT* qv = reinterpret_cast<const VT*>(sq);
  const VT* src = reinterpret_cast<const VT*>(rows + (size_t)base * sdim);
  if ((sdimv & 31) == 0) {
    // every lane always handles the same chunks of the query
    for (int c = lane; c < sdimv; c += 32) {
      const VT qc = qv[c];
#pragma unroll 4
      for (int r = 0; r < rows_here; ++r)
        tile_v[r * pitch + c] = term_vec<METRIC>(qc, ldg16(src + (size_t)r * sdimv + c));
    }
  } else {
    const int total = rows_here * sdimv;
#pragma unroll 2
    for (int t = lane; t < total; t += 32) {
      const int r = t / sdimv, c = t - r * sdimv;
      tile_v[r * pitch + c] = term_vec<METRIC>(qv[c], ldg16(src + t));
    }
  }
  __syncwarp();
  T d = metric_init<T>(METRIC);
  if (lane < rows_here) {
    const VT* row = tile_v + lane * pitch;
#pragma unroll 8
    for (int c = 0; c < sdimv; ++c) d = acc_vec<METRIC>(d, row[c]);
  }
  __syncwarp();  // the tile is overwritten by the next round
  return d;
}

// Returns, in lane p < rows_here (rows_here <= tile_rows), the distance of row base + p to the query
// (scaled for the approximate visitors) and its original index.
template <typename T>
__device__ __forceinline__ void stage_and_fold(const T* __restrict__ rows, const int32_t* __restrict__ indices, int sdim,
                                               int base, int rows_here, const T* sq, T* tile, int metric, bool approx,
                                               T e_inv, T& d, int& idx) {
  const int lane = threadIdx.x & 31;
  const bool valid = lane < rows_here;
  idx = -1;
  if (valid) idx = __ldg(indices + base + lane);
  switch (metric) {
    case PICO_B200_METRIC_L2_SQUARED:
      d = terms_then_fold<PICO_B200_METRIC_L2_SQUARED, T>(rows, sdim, base, rows_here, sq, tile);
      break;
    case PICO_B200_METRIC_L1:
      d = terms_then_fold<PICO_B200_METRIC_L1, T>(rows, sdim, base, rows_here, sq, tile);
      break;
    case PICO_B200_METRIC_LPINF:
      d = terms_then_fold<PICO_B200_METRIC_LPINF, T>(rows, sdim, base, rows_here, sq, tile);
      break;
    default:
      d = terms_then_fold<PICO_B200_METRIC_LNINF, T>(rows, sdim, base, rows_here, sq, tile);
      break;
  }
  if (!valid)
    d = Limits<T>::max();
  else if (approx)
    d = mul_rn(d, e_inv);
}

template <typename T, typename Visitor>
__device__ __forceinline__ void scan_leaf_staged(const T* __restrict__ rows, const int32_t* __restrict__ indices,
                                                 int sdim, int lb, int le, const T* sq, T* tile, int tile_rows,
                                                 int metric, bool approx, T e_inv, Visitor& vis) {
  const int lane = threadIdx.x & 31;
  for (int base = lb; base < le; base += tile_rows) {
    const int rows_here = min(tile_rows, le - base);
    T d;
    int idx;
    stage_and_fold<T>(rows, indices, sdim, base, rows_here, sq, tile, metric, approx, e_inv, d, idx);
    vis.visit_batch(lane < rows_here, idx, d);
  }
}

// Subtree distance cache (row storage). Leaves of a high-dimensional tree are small (2-3 points on
// average under the sliding midpoint rule with max_leaf_size 10), so a lane-per-point leaf scan
// keeps 2 of 32 lanes busy and the kernel is bound by instruction issue. Points are stored in
// leaf order, i.e. the points below ANY node form one contiguous run: when the descent reaches a
// node with at most cache_rows points below it, the distances of all of them are computed in
// full-tile rounds (one lane per point) and kept in shared memory; the traversal below that node —
// its order, its prune tests, what is offered to the visitor — runs unchanged and takes the
// distances from the cache. Distances of leaves that end up pruned were computed for nothing; the
// lanes would have idled anyway.
template <typename T>
struct SpanCache {
  T* dist;        // [cache_rows] shared
  int* index;     // [cache_rows] shared
  int lo, hi;     // cached point range
  int base_sp;    // stack height when it was filled; frames below it belong to ancestors
};

// ---------------------------------------------------------------- warp visitors
// k <= 32: the sorted list lives in registers, slot i in lane i ("register k-heap").
template <typename T>
struct WarpKnnReg {
  T d;
  int id;
  T worst;
  int k;
  __device__ __forceinline__ void init(int k_) {
    k = k_;
    d = Limits<T>::max();
    id = -1;
    worst = Limits<T>::max();
  }
  __device__ __forceinline__ T max() const { return worst; }
  // insert_sorted (search_visitor.hpp:24-38): new item goes behind all items <= it.
  __device__ __forceinline__ void insert(int i_new, T x) {
    const int lane = threadIdx.x & 31;
    const unsigned le = __ballot_sync(0xffffffffu, lane < k && d <= x);
    const int p = __popc(le);
    const T up_d = __shfl_up_sync(0xffffffffu, d, 1);
    const int up_i = __shfl_up_sync(0xffffffffu, id, 1);
    if (lane == p) {
      d = x;
      id = i_new;
    } else if (lane > p) {
      d = up_d;
      id = up_i;
    }
    worst = __shfl_sync(0xffffffffu, d, k - 1);
  }
  __device__ __forceinline__ void store(Neighbor<T>* out) const {
    const int lane = threadIdx.x & 31;
    if (lane < k) {
      out[lane].index = id;
      out[lane].distance = d;
    }
  }
};

// any k: the output row itself is the sorted working list, like the reference's iterator
// range (search_visitor.hpp:98-123).
template <typename T>
struct WarpKnnMem {
  Neighbor<T>* list;
  int k, active;
  T worst;
  __device__ __forceinline__ void init(Neighbor<T>* row, int k_) {
    list = row;
    k = k_;
    active = 0;
    worst = Limits<T>::max();
    const int lane = threadIdx.x & 31;
    for (int i = lane; i < k; i += 32) {
      list[i].index = -1;
      list[i].distance = Limits<T>::max();
    }
    __syncwarp();
  }
  __device__ __forceinline__ T max() const { return worst; }
  __device__ __forceinline__ void insert(int i_new, T x) {
    const int lane = threadIdx.x & 31;
    if (active < k) ++active;
    int c = 0;
    for (int i = lane; i < active - 1; i += 32) c += (list[i].distance <= x);
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    const int p = c;
    for (int hi = active - 1; hi > p; hi -= 32) {
      const int i = hi - lane;
      Neighbor<T> v;
      if (i > p) v = list[i - 1];
      __syncwarp();
      if (i > p) list[i] = v;
      __syncwarp();
    }
    if (lane == 0) {
      list[p].index = i_new;
      list[p].distance = x;
    }
    __syncwarp();
    worst = (active == k) ? list[k - 1].distance : Limits<T>::max();
  }
  __device__ __forceinline__ void store(Neighbor<T>*) const {}
};

template <typename T, typename List>
struct WarpVisitKnn {
  List list;
  __device__ __forceinline__ T max() const { return list.max(); }
  __device__ __forceinline__ void visit_batch(bool valid, int idx, T d) {
    unsigned cand = __ballot_sync(0xffffffffu, valid && list.max() > d);
    while (cand) {
      const int l = __ffs(cand) - 1;
      cand &= cand - 1;
      const T x = __shfl_sync(0xffffffffu, d, l);
      const int xi = __shfl_sync(0xffffffffu, idx, l);
      if (list.max() > x) list.insert(xi, x);  // max() may have shrunk since the ballot
    }
  }
};

template <typename T>
struct WarpVisitRadius {
  T radius;
  Neighbor<T>* out;  // nullptr while counting
  uint32_t count = 0;
  __device__ __forceinline__ T max() const { return radius; }
  __device__ __forceinline__ void visit_batch(bool valid, int idx, T d) {
    const unsigned hits = __ballot_sync(0xffffffffu, valid && radius > d);
    if (out) {
      const int lane = threadIdx.x & 31;
      if (hits & (1u << lane)) {
        Neighbor<T>* o = out + count + __popc(hits & ((1u << lane) - 1u));
        o->index = idx;
        o->distance = d;
      }
    }
    count += __popc(hits);
  }
};

// ---------------------------------------------------------------- nearest traversal
// `sq` = query, `so` = node_box_offset_ (both shared, sdim entries, warp-private).
template <typename T, bool PACKED, typename Visitor>
__device__ void traverse_warp(const typename NodeOf<T>::type* __restrict__ nodes, const T* __restrict__ outer,
                              const PointSet<T, PACKED>& ps, const T* sq, T* so, WarpFrame<T>* stack,
                              WarpFrame<T>* win, int metric, bool approx, T e_inv, Visitor& vis, T* tile = nullptr,
                              int tile_rows = 0, const uint2* __restrict__ spans = nullptr, int cache_rows = 0) {
  // `stack` (global, one slot per tree level) is the backing store; `win` (shared, kStackWindow
  // frames) mirrors the most recently pushed ones, so the pop that follows a push — every leaf
  // visit — does not wait for a global-memory round trip. Frames [win_lo, sp) are valid in `win`.
  const int lane = threadIdx.x & 31;
  uint32_t node = 0;
  T node_dist = T(0);
  int sp = 0;
  int win_lo = 0;
  SpanCache<T> cache;
  cache.lo = cache.hi = 0;
  cache.base_sp = 0x7fffffff;  // nothing cached
  if (!PACKED && tile_rows > 0) {
    // behind the tile: cache_rows distances, then cache_rows indices
    cache.dist = tile + (size_t)tile_rows * (ps.sdim + Vec16<T>::n);
    cache.index = reinterpret_cast<int*>(cache.dist + cache_rows);
  }
  for (;;) {
    T a, b;
    uint32_t right, sd;
    int lb, le;
    load_node(nodes, node, a, b, right, sd, lb, le);
    while (sd != PICO_B200_LEAF) {
      if (!PACKED && spans != nullptr && sp < cache.base_sp) {
        // outside any cached subtree: does everything below this node fit into one tile?
        const uint2 span = __ldg(spans + node);
        if ((int)span.y <= cache_rows) {
          // rounds of tile_rows rows: every round is a full tile although the leaves are small
          for (int done = 0; done < (int)span.y; done += tile_rows) {
            const int rows_here = min(tile_rows, (int)span.y - done);
            T d;
            int idx;
            stage_and_fold<T>(ps.rows, ps.indices, ps.sdim, (int)span.x + done, rows_here, sq, tile, metric, approx,
                              e_inv, d, idx);
            if (lane < rows_here) {
              cache.dist[done + lane] = d;
              cache.index[done + lane] = idx;
            }
          }
          __syncwarp();
          cache.lo = (int)span.x;
          cache.hi = (int)(span.x + span.y);
          cache.base_sp = sp;
        }
      }
      const T v = sq[sd];
      bool go_left;
      T new_off;
      branch_choice(metric, outer, node, a, b, v, sd, go_left, new_off);
      const uint32_t far = go_left ? right : node + 1;
      if (lane == 0) {
        WarpFrame<T> f;
        f.far = far;
        f.sd_state = sd;
        f.dist = node_dist;
        f.new_off = new_off;
        f.old_off = T(0);
        win[sp & (kStackWindow - 1)] = f;
        stack[sp] = f;
        prefetch_l2(nodes + far);  // it is (most likely) visited after the near subtree
      }
      win_lo = max(win_lo, sp + 1 - kStackWindow);
      ++sp;
      node = go_left ? node + 1 : right;
      load_node(nodes, node, a, b, right, sd, lb, le);
    }
    if (!PACKED && tile_rows > 0 && sp >= cache.base_sp && lb >= cache.lo && le <= cache.hi) {
      for (int base = lb; base < le; base += 32) {
        const int i = base + lane;
        const bool valid = i < le;
        const T d = valid ? cache.dist[i - cache.lo] : Limits<T>::max();
        const int idx = valid ? cache.index[i - cache.lo] : -1;
        vis.visit_batch(valid, idx, d);
      }
    } else if (!PACKED && tile_rows > 0) {
      scan_leaf_staged<T>(ps.rows, ps.indices, ps.sdim, lb, le, sq, tile, tile_rows, metric, approx, e_inv, vis);
    } else {
      for (int base = lb; base < le; base += 32) {
        const int i = base + lane;
        const bool valid = i < le;
        int idx = -1;
        T d = Limits<T>::max();
        if (valid) {
          d = ps.distance(i, sq, metric, idx);
          if (approx) d = mul_rn(d, e_inv);
        }
        vis.visit_batch(valid, idx, d);
      }
    }
    // unwind (kd_tree_search.hpp:93-103)
    bool found = false;
    __syncwarp();
    while (sp > 0) {
      const int top = sp - 1;
      WarpFrame<T> f;
      if (top >= win_lo) {
        f = win[top & (kStackWindow - 1)];
      } else {
        f = stack[top];
        __syncwarp();
        if (lane == 0) win[top & (kStackWindow - 1)] = f;
        win_lo = top;
        __syncwarp();
      }
      const uint32_t fsd = f.sd_state & 0x7fffffffu;
      if (!(f.sd_state >> 31)) {
        const T old = so[fsd];
        const T d2 = add_rn(sub_rn(f.dist, old), f.new_off);
        if (vis.max() >= d2) {
          __syncwarp();
          if (lane == 0) {
            WarpFrame<T>& wf = win[top & (kStackWindow - 1)];
            wf.sd_state = fsd | 0x80000000u;
            wf.old_off = old;
            stack[top].sd_state = fsd | 0x80000000u;
            stack[top].old_off = old;
            so[fsd] = f.new_off;
          }
          __syncwarp();
          node = f.far;
          node_dist = d2;
          found = true;
          if (top < cache.base_sp) cache.base_sp = 0x7fffffff;  // left the cached subtree
          break;
        }
        --sp;
      } else {
        __syncwarp();
        if (lane == 0) so[fsd] = f.old_off;
        __syncwarp();
        --sp;
      }
    }
    if (!found) return;
  }
}

// ---------------------------------------------------------------- box traversal
struct BoxFrame {
  uint32_t node;
  uint32_t stage;
};

// search_box::operator() (kd_tree_search.hpp:270-306). `sbox` = running box_ (min then
// max), `qmin/qmax` = query box, all shared and warp-private; `saved` holds the value a
// frame has to put back. With out == nullptr only counts.
template <typename T, bool PACKED>
__device__ uint32_t traverse_box_warp(const typename NodeOf<T>::type* __restrict__ nodes,
                                      const T* __restrict__ outer, int metric, const PointSet<T, PACKED>& ps,
                                      const int32_t* __restrict__ indices, const T* qmin, const T* qmax, T* sbox,
                                      BoxFrame* stack, T* saved, int32_t* out) {
  const int lane = threadIdx.x & 31;
  const int sdim = ps.sdim;
  uint32_t count = 0;
  int sp = 0;
  if (lane == 0) {
    stack[0].node = 0;
    stack[0].stage = 0;
  }
  sp = 1;
  __syncwarp();
  while (sp > 0) {
    const BoxFrame f = stack[sp - 1];
    T a, b;
    uint32_t right, sd;
    int lb, le;
    load_node(nodes, f.node, a, b, right, sd, lb, le);
    if (sd == PICO_B200_LEAF) {
      for (int base = lb; base < le; base += 32) {
        const int i = base + lane;
        int idx = -1;
        const bool in = (i < le) && ps.inside(i, qmin, qmax, metric, idx);
        const unsigned hits = __ballot_sync(0xffffffffu, in);
        if (out && in) out[count + __popc(hits & ((1u << lane) - 1u))] = idx;
        count += __popc(hits);
      }
      --sp;
      continue;
    }
    if (f.stage == 2) {
      __syncwarp();
      if (lane == 0) sbox[sd] = saved[sp - 1];
      __syncwarp();
      --sp;
      continue;
    }
    // stage 0: narrow max to left_max and look at the left child;
    // stage 1: put max back, narrow min to right_min and look at the right child.
    const uint32_t child = (f.stage == 0) ? f.node + 1 : right;
    __syncwarp();
    if (lane == 0) {
      if (f.stage == 0) {
        saved[sp - 1] = sbox[sdim + sd];
        sbox[sdim + sd] = a;
      } else {
        sbox[sdim + sd] = saved[sp - 1];
        saved[sp - 1] = sbox[sd];
        sbox[sd] = b;
      }
      stack[sp - 1].stage = f.stage + 1;
    }
    __syncwarp();
    // query.contains(box_) := contains(box.min) && contains(box.max)  (box.hpp:44-47); topological:
    // metric_box_map::contains(box), box.hpp:317-329 — per dimension the cell [min, max] must lie in
    // the query segment, which on the circle may wrap (segment_s1::contains(segment_r1), segment.hpp:71-77)
    bool contained = true;
    if (!is_topological(metric)) {
      for (int j = lane; j < 2 * sdim; j += 32) {
        const int dj = j < sdim ? j : j - sdim;
        const T x = sbox[j];
        if (qmin[dj] > x || qmax[dj] < x) contained = false;
      }
    } else {
      for (int j = lane; j < sdim; j += 32) {
        const T mn = qmin[j], mx = qmax[j], cmin = sbox[j], cmax = sbox[sdim + j];
        bool in;
        if (!dim_is_s1(metric, (uint32_t)j) || mn <= mx)
          in = mn <= cmin && cmax <= mx;
        else
          in = cmin >= mn || cmax <= mx;
        if (!in) contained = false;
      }
    }
    contained = __all_sync(0xffffffffu, contained);
    if (contained) {
      // report_node (kd_tree_search.hpp:336-372): the subtree's points are one contiguous run
      uint32_t nl = child, nr = child;
      int rb, re, dummy0, dummy1;
      T ta, tb;
      uint32_t tr, tsd;
      load_node(nodes, nl, ta, tb, tr, tsd, rb, dummy0);
      while (tsd != PICO_B200_LEAF) {
        ++nl;
        load_node(nodes, nl, ta, tb, tr, tsd, rb, dummy0);
      }
      load_node(nodes, nr, ta, tb, tr, tsd, dummy1, re);
      while (tsd != PICO_B200_LEAF) {
        nr = tr;
        load_node(nodes, nr, ta, tb, tr, tsd, dummy1, re);
      }
      if (out)
        for (int i = rb + lane; i < re; i += 32) out[count + (i - rb)] = __ldg(indices + i);
      count += (uint32_t)(re - rb);
    } else {
      // intersects_left / intersects_right, kd_tree_search.hpp:310-328
      bool intersects = (f.stage == 0) ? (qmin[sd] <= a) : (qmax[sd] >= b);
      if (is_topological(metric)) {
        T lmin, rmax;
        load_outer(outer, f.node, lmin, rmax);
        intersects = intersects || ((f.stage == 0) ? (qmax[sd] >= lmin) : (qmin[sd] <= rmax));
      }
      if (intersects) {
        if (lane == 0) {
          stack[sp].node = child;
          stack[sp].stage = 0;
        }
        ++sp;
        __syncwarp();
      }
    }
  }
  return count;
}

}  // namespace pico
