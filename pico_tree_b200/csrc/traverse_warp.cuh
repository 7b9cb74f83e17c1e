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

// ---------------------------------------------------------------- staged leaf scan (row storage)
// sdim > 3: a leaf is `rows x sdim` contiguous scalars. With one lane per point reading its own
// row, every load instruction touches as many sectors as there are points and only ~10 of 32
// lanes issue loads at all (profiles/r1/configs_v3.jsonl: 128-D exact search ran at 350 GB/s).
// Instead the whole warp copies the leaf tile to shared memory with 16-byte cp.async (fully
// coalesced, no registers), then lane p folds row p in dimension order — the summation order, and
// with it every distance bit, stays that of metric.hpp:36-51.
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

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

__device__ __forceinline__ float fold_vec(int metric, float d, const float4& q, const float4& p, int j) {
  d = metric_fold(metric, d, q.x, p.x, j);
  d = metric_fold(metric, d, q.y, p.y, j + 1);
  d = metric_fold(metric, d, q.z, p.z, j + 2);
  return metric_fold(metric, d, q.w, p.w, j + 3);
}
__device__ __forceinline__ double fold_vec(int metric, double d, const double2& q, const double2& p, int j) {
  d = metric_fold(metric, d, q.x, p.x, j);
  return metric_fold(metric, d, q.y, p.y, j + 1);
}

// One row against the query, dimension order, with the metric resolved OUTSIDE the loop (a
// run-time switch per coordinate costs more than the arithmetic).
template <int METRIC, typename T, typename VT>
__device__ __forceinline__ T fold_row(const VT* __restrict__ qv, const VT* __restrict__ row, int sdimv) {
  constexpr int V = Vec16<T>::n;
  T d = metric_init<T>(METRIC);
#pragma unroll 8
  for (int c = 0; c < sdimv; ++c) d = fold_vec(METRIC, d, qv[c], row[c], c * V);
  return d;
}

// tile: tile_rows x (sdim + Vec16::n) scalars of warp-private shared memory, 16-byte aligned;
// sq: the query, same alignment. Requires sdim % Vec16::n == 0.
template <typename T, typename Visitor>
__device__ __forceinline__ void scan_leaf_staged(const T* __restrict__ rows, const int32_t* __restrict__ indices,
                                                 int sdim, int lb, int le, const T* sq, T* tile, int tile_rows,
                                                 int metric, bool approx, T e_inv, Visitor& vis) {
  using VT = typename Vec16<T>::type;
  constexpr int V = Vec16<T>::n;
  const int lane = threadIdx.x & 31;
  const int sdimv = sdim / V;
  const int pitch = sdimv + 1;  // one 16-byte pad per row: rows start in different banks
  VT* tile_v = reinterpret_cast<VT*>(tile);
  const VT* qv = reinterpret_cast<const VT*>(sq);
  for (int base = lb; base < le; base += tile_rows) {
    const int rows_here = min(tile_rows, le - base);
    const VT* src = reinterpret_cast<const VT*>(rows + (size_t)base * sdim);
    if ((sdimv & 31) == 0) {
      for (int r = 0; r < rows_here; ++r)
        for (int c = lane; c < sdimv; c += 32) cp_async16(tile_v + r * pitch + c, src + (size_t)r * sdimv + c);
    } else {
      const int total = rows_here * sdimv;
      for (int t = lane; t < total; t += 32) {
        const int r = t / sdimv, c = t - r * sdimv;
        cp_async16(tile_v + r * pitch + c, src + t);
      }
    }
    const bool valid = lane < rows_here;
    int idx = -1;
    if (valid) idx = __ldg(indices + base + lane);  // in flight together with the tile
    cp_async_wait_all();
    __syncwarp();
    T d = Limits<T>::max();
    if (valid) {
      const VT* row = tile_v + lane * pitch;
      switch (metric) {
        case PICO_B200_METRIC_L2_SQUARED:
          d = fold_row<PICO_B200_METRIC_L2_SQUARED, T>(qv, row, sdimv);
          break;
        case PICO_B200_METRIC_L1:
          d = fold_row<PICO_B200_METRIC_L1, T>(qv, row, sdimv);
          break;
        case PICO_B200_METRIC_LPINF:
          d = fold_row<PICO_B200_METRIC_LPINF, T>(qv, row, sdimv);
          break;
        default:
          d = fold_row<PICO_B200_METRIC_LNINF, T>(qv, row, sdimv);
          break;
      }
      if (approx) d = mul_rn(d, e_inv);
    }
    __syncwarp();  // the tile is overwritten by the next round
    vis.visit_batch(valid, idx, d);
  }
}

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
                              int tile_rows = 0) {
  // `stack` (global, one slot per tree level) is the backing store; `win` (shared, kStackWindow
  // frames) mirrors the most recently pushed ones, so the pop that follows a push — every leaf
  // visit — does not wait for a global-memory round trip. Frames [win_lo, sp) are valid in `win`.
  const int lane = threadIdx.x & 31;
  uint32_t node = 0;
  T node_dist = T(0);
  int sp = 0;
  int win_lo = 0;
  for (;;) {
    T a, b;
    uint32_t right, sd;
    int lb, le;
    load_node(nodes, node, a, b, right, sd, lb, le);
    while (sd != PICO_B200_LEAF) {
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
    if (!PACKED && tile_rows > 0) {
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
