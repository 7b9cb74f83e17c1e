/*
 * pico_b200.h — C-ABI of libpico_b200.so, the B200 (sm_100a) KdTree engine that
 * replaces PicoTree's build + nearest/radius/box search internals.
 *
 * The reference (Jaybro/pico_tree v1.0.0) is a header-only C++ template library
 * with no FFI of its own. The boundary therefore sits where its public class
 * hands over to its `internal::` algorithms; each entry point below names the
 * reference code it stands in for (paths relative to src/pico_tree/pico_tree/
 * unless they start with src/).
 *
 * Conventions
 *   - every function returns 0 on success or a pico_b200_status code;
 *     pico_b200_last_error() gives a thread-local message for the last failure.
 *   - plain pointers + sizes only. Unless a flag says otherwise pointers are
 *     HOST pointers and the call copies host<->device itself.
 *   - points / queries are row-major, `stride` counted in scalars (>= sdim).
 *   - neighbours are written as pico_tree::neighbor<int, Scalar> records
 *     (core.hpp:24-46): {int32 index; Scalar distance} = 8 B (f32) / 16 B (f64).
 *   - a tree handle is immutable after creation; searches on one handle may be
 *     issued from several host threads (each call uses its own stream and
 *     workspace), matching the reference's "thread safe queries".
 *   - there is no CPU fallback: without a CUDA device every compute entry point
 *     fails with PICO_B200_ERR_NO_DEVICE.
 */
#ifndef PICO_B200_H_
#define PICO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PICO_B200_ABI_VERSION 1

typedef enum pico_b200_status {
  PICO_B200_OK = 0,
  PICO_B200_ERR_INVALID_ARGUMENT = 1,
  PICO_B200_ERR_NO_DEVICE = 2,
  PICO_B200_ERR_CUDA = 3,
  PICO_B200_ERR_UNSUPPORTED = 4,
  PICO_B200_ERR_OUT_OF_MEMORY = 5,
  PICO_B200_ERR_NCCL = 6
} pico_b200_status;

/* Scalar type of the space (space_traits<>::scalar_type). */
typedef enum { PICO_B200_F32 = 0, PICO_B200_F64 = 1 } pico_b200_scalar;

/* metric.hpp:77-190. Values follow the order used by the oracle. */
typedef enum {
  PICO_B200_METRIC_L1 = 0,         /* metric_l1          metric.hpp:77-97   */
  PICO_B200_METRIC_L2_SQUARED = 1, /* metric_l2_squared  metric.hpp:99-123  */
  PICO_B200_METRIC_LPINF = 2,      /* metric_lpinf       metric.hpp:125-151 */
  PICO_B200_METRIC_LNINF = 3,      /* metric_lninf       metric.hpp:153-180 */
  /* topological spaces (search_nearest_topological, internal/kd_tree_search.hpp:122-229): */
  PICO_B200_METRIC_SO2 = 4,        /* metric_so2         metric.hpp:197-221, sdim 1: S1 = [0,1)      */
  PICO_B200_METRIC_SE2_SQUARED = 5,/* metric_se2_squared metric.hpp:223-257, sdim 3: R2 x S1         */
  /* A user-defined Metric_ (kd_tree.hpp:19-36 accepts any type; examples/kd_tree/kd_tree_custom_metric.cpp): the
   * tree is built, exported, saved and loaded here — the topological flavour keeps all four bounds per node —
   * but knn / radius need the caller's functor and are answered by the host header
   * (include/pico_tree_b200/host_search.hpp); pico_b200_knn / pico_b200_radius return PICO_B200_ERR_UNSUPPORTED.
   * pico_b200_box works for the euclidean flavour (a box test needs no metric). */
  PICO_B200_METRIC_CUSTOM_TOPOLOGICAL = 6,
  PICO_B200_METRIC_CUSTOM_EUCLIDEAN = 7
} pico_b200_metric;

/* internal/kd_tree_builder.hpp:35-75 */
typedef enum {
  PICO_B200_RULE_SLIDING_MIDPOINT_MAX_SIDE = 0,
  PICO_B200_RULE_MIDPOINT_MAX_SIDE = 1,
  PICO_B200_RULE_MEDIAN_MAX_SIDE = 2
} pico_b200_rule;

/* internal/kd_tree_builder.hpp:92-104 */
typedef enum { PICO_B200_STOP_MAX_LEAF_SIZE = 0, PICO_B200_STOP_MAX_LEAF_DEPTH = 1 } pico_b200_stop;

/* Flags for the search calls. */
enum {
  PICO_B200_DEVICE_POINTERS = 1u << 0, /* query/result pointers are device pointers on the tree's device */
  PICO_B200_NO_REORDER = 1u << 1,      /* do not Z-order the query batch before traversal            */
  PICO_B200_SORT_RESULTS = 1u << 2,    /* search_radius(..., sort = true) kd_tree.hpp:265-267        */
  PICO_B200_WARP_PER_QUERY = 1u << 3,  /* force the warp-per-query traversal kernel                  */
  PICO_B200_ASYNC = 1u << 4            /* with DEVICE_POINTERS + pico_b200_set_stream: enqueue only,  */
                                       /* do not synchronise (stats are not filled)                   */
};

typedef struct pico_b200_tree pico_b200_tree;

/*
 * Flat pre-order node, the on-device replacement of kd_tree_node_euclidean
 * (internal/kd_tree_node.hpp:80-96) and its pointer links. Node 0 is the root,
 * the left child of branch i is node i+1, the right child is `right`.
 *   branch: a = left_max, b = right_min (tight child bounds on split_dim, as
 *           set_branch stores them, kd_tree_node.hpp:86-92), split_dim < sdim
 *   leaf:   a = begin_idx, b = end_idx (bit patterns of int32),
 *           right = 0xFFFFFFFF, split_dim = 0xFFFFFFFF
 * The f64 flavour widens a and b to 8 bytes (24 B padded to 32 B).
 */
typedef struct pico_b200_node_f32 {
  union { float left_max; int32_t begin_idx; } a;
  union { float right_min; int32_t end_idx; } b;
  uint32_t right;
  uint32_t split_dim;
} pico_b200_node_f32;

typedef struct pico_b200_node_f64 {
  union { double left_max; int64_t begin_idx; } a;
  union { double right_min; int64_t end_idx; } b;
  uint32_t right;
  uint32_t split_dim;
  uint64_t pad_;
} pico_b200_node_f64;

#define PICO_B200_LEAF 0xFFFFFFFFu

typedef struct pico_b200_tree_info {
  uint64_t n_points, sdim, n_nodes, n_leaves, height;
  int32_t scalar, metric, device, reserved_;
  double build_ms;          /* device time of the build kernels (CUDA events)          */
  uint64_t device_bytes;    /* HBM held by the handle                                  */
} pico_b200_tree_info;

/* Per-call device timing, filled when a non-NULL pointer is passed. */
typedef struct pico_b200_search_stats {
  double h2d_ms, reorder_ms, kernel_ms, d2h_ms;
  uint64_t kernel_launches;
} pico_b200_search_stats;

const char* pico_b200_last_error(void);
int pico_b200_abi_version(void);
int pico_b200_device_count(int* count);

/*
 * Builds the tree on `device`.
 * Replaces internal::build_kd_tree::operator() (internal/kd_tree_builder.hpp:471-494),
 * build_kd_tree_impl::create_node (:351-396), the three splitters (:143-280) and
 * space_wrapper::compute_bounding_box (internal/space_wrapper.hpp:34-40), i.e. what the
 * kd_tree constructor runs (kd_tree.hpp:76-88).
 *   bounds_min/max: NULL = bounds_from_space (:120-125), else bounds_t<P> (:127-138).
 * The point set is copied to the device (the reference keeps a view; mutating the host
 * array afterwards already invalidates its index too).
 */
int pico_b200_tree_create(const void* pts, size_t n, size_t sdim, size_t stride, int scalar, int metric,
                          int rule, int stop_kind, size_t stop_value, const void* bounds_min,
                          const void* bounds_max, int device, pico_b200_tree** out);

/*
 * Uploads an existing tree (pre-order nodes + index permutation + root box).
 * Replaces kd_tree_data::load / read_node (internal/kd_tree_data.hpp:43-52,89-107), i.e.
 * kd_tree::load (kd_tree.hpp:336-353).
 *   outer_bounds: NULL for the euclidean metrics. For metric_so2 / metric_se2_squared the nodes
 *   are kd_tree_node_topological (internal/kd_tree_node.hpp:99-117) with four bounds per branch;
 *   left_max / right_min live in the node record, {left_min, right_max} come as n_nodes pairs here.
 */
int pico_b200_tree_create_from_nodes(const void* pts, size_t n, size_t sdim, size_t stride, int scalar,
                                     int metric, const void* nodes, size_t n_nodes, const int32_t* indices,
                                     const void* root_box_min_then_max, const void* outer_bounds, int device,
                                     pico_b200_tree** out);

void pico_b200_tree_destroy(pico_b200_tree* tree);

int pico_b200_tree_info_get(const pico_b200_tree* tree, pico_b200_tree_info* info);

/*
 * Copies the flat tree back to the host: `nodes_out` (n_nodes records of the
 * matching pico_b200_node_*), `indices_out` (n int32: kd_tree_data::indices,
 * internal/kd_tree_data.hpp:64) and `root_box_out` (min[sdim] then max[sdim]).
 * Any pointer may be NULL. Serves kd_tree::save (kd_tree.hpp:355-370) and
 * kd_tree::leaf_ranges (kd_tree.hpp:325, kd_tree_data.hpp:60-62,76-87).
 */
int pico_b200_tree_export(const pico_b200_tree* tree, void* nodes_out, int32_t* indices_out, void* root_box_out);
/* topological metrics only: n_nodes pairs {left_min, right_max} (zeros for leaves) */
int pico_b200_tree_export_outer_bounds(const pico_b200_tree* tree, void* outer_out);

/*
 * k nearest neighbours for a batch of queries; out holds nq*k neighbour records,
 * row i sorted ascending. k is NOT clamped here: callers clamp to n like
 * kd_tree.hpp:193 does. e <= 0 selects the exact visitors, e > 0 the approximate
 * ones (returned distances are scaled by 1/e, search_visitor.hpp:178-183,226-237).
 * Replaces search_nearest_euclidean (internal/kd_tree_search.hpp:24-113) driven by
 * the search_nn / search_knn / search_approximate_* visitors
 * (internal/search_visitor.hpp:41-123,164-247), and the batch loop of the binding
 * (src/pyco_tree/pico_tree/_pyco_tree/kd_tree.hpp:117-135,143-165) and of the benchmark
 * (examples/benchmark/bm_pico_kd_tree.cpp:63-78).
 */
int pico_b200_knn(const pico_b200_tree* tree, const void* queries, size_t nq, size_t stride, size_t k,
                  double e, void* neighbors_out, unsigned flags, pico_b200_search_stats* stats);

/*
 * All neighbours with distance < radius (strict), ragged: offsets_out[nq+1] and one
 * malloc'd array of neighbour records (free with pico_b200_free). Visit order unless
 * PICO_B200_SORT_RESULTS. With PICO_B200_DEVICE_POINTERS `queries` and `offsets_out` are device
 * pointers and *neighbors_out receives a device buffer (pico_b200_free_device): results stay in
 * HBM for a consumer kernel. Replaces kd_tree::search_radius (kd_tree.hpp:256-290) with
 * the search_radius / search_approximate_radius visitors
 * (internal/search_visitor.hpp:126-156,253-288) and the binding loop
 * (_pyco_tree/kd_tree.hpp:170-238).
 */
int pico_b200_radius(const pico_b200_tree* tree, const void* queries, size_t nq, size_t stride, double radius,
                     double e, uint64_t* offsets_out, void** neighbors_out, unsigned flags,
                     pico_b200_search_stats* stats);

/*
 * Axis-aligned box queries (inclusive bounds), ragged int32 indices in DFS order.
 * Replaces internal::search_box (internal/kd_tree_search.hpp:237-381) behind
 * kd_tree::search_box (kd_tree.hpp:295-318) and the binding loop
 * (_pyco_tree/kd_tree.hpp:240-268).
 */
int pico_b200_box(const pico_b200_tree* tree, const void* mins, const void* maxs, size_t nb, size_t stride,
                  uint64_t* offsets_out, int32_t** indices_out, unsigned flags, pico_b200_search_stats* stats);

/*
 * Multi-GPU replicas: one process per GPU, tree built on `root`, broadcast over
 * NVLink with NCCL (the reference has no distributed path).
 * `nccl_comm` is an ncclComm_t. On non-root ranks `*tree` may be NULL on entry
 * and receives a new handle on `device`.
 */
int pico_b200_tree_broadcast(pico_b200_tree** tree, void* nccl_comm, int rank, int root, int device);

/* Serialised image of a handle, for callers that move trees with their own
 * transport (torch.distributed broadcast of a byte tensor). */
int pico_b200_tree_serialize_size(const pico_b200_tree* tree, uint64_t* bytes);
int pico_b200_tree_serialize(const pico_b200_tree* tree, void* dst, int dst_is_device);
int pico_b200_tree_deserialize(const void* src, uint64_t bytes, int src_is_device, int device,
                               pico_b200_tree** out);

/*
 * The reference's own on-disk tree image, byte for byte: what kd_tree::save writes and
 * kd_tree::load reads (kd_tree.hpp:336-370 -> kd_tree_data::save/load/read/write and
 * write_node/read_node, internal/kd_tree_data.hpp:43-58,89-135): size_t sdim; size_t n;
 * int32 indices[n]; Scalar min[sdim]; Scalar max[sdim]; then the nodes in pre-order, each a
 * 1-byte is_leaf flag followed by the raw leaf {int32 begin, end} or the raw branch
 * {int32 split_dim; Scalar left_max; Scalar right_min} (12 B for f32, 24 B for f64); topological metrics:
 * {int32 split_dim; Scalar left_min, left_max, right_min, right_max} (20 B / 40 B, kd_tree_node.hpp:52-59).
 * Trees written by either side load in the other. The points are not part of the image.
 *   pico_b200_tree_load: `consumed` (may be NULL) receives the number of bytes read, so a
 *   caller can keep reading its own trailer/header around the image (the .pkd header of
 *   src/pyco_tree/pico_tree/_pyco_tree/kd_tree.hpp:547-614 is written by the host side).
 */
int pico_b200_tree_save_size(const pico_b200_tree* tree, uint64_t* bytes);
int pico_b200_tree_save(const pico_b200_tree* tree, void* dst);
int pico_b200_tree_load(const void* pts, size_t n, size_t sdim, size_t stride, int scalar, int metric,
                        const void* stream, uint64_t stream_bytes, int device, pico_b200_tree** out,
                        uint64_t* consumed);

/*
 * ---- kd_forest ---------------------------------------------------------------
 * Replaces pico_tree::kd_forest (examples/pico_understory/pico_understory/kd_forest.hpp:15-138):
 * `forest_size` kd-trees (sliding_midpoint_max_side, max_leaf_size, bounds from the space) over
 * Householder-reflected copies of the point set (internal/rkd_tree_hh_data.hpp:51-90,
 * internal/rkd_tree_builder.hpp:26-41), searched best-bin-first with at most `max_leaves_visited`
 * leaves per tree and one neighbour list shared by all trees
 * (internal/kd_tree_priority_search.hpp:24-142, kd_forest.hpp:91-120). metric_l2_squared only, like
 * the reference's priority search. Distances are measured in each tree's reflected space, as in the
 * reference.
 *   rotations  forest_size unit vectors of sdim scalars (row-major), or NULL: drawn at random like
 *              rkd_tree_hh_data::random_rotation (std::random_device-seeded, not reproducible)
 */
typedef struct pico_b200_forest pico_b200_forest;

typedef struct pico_b200_forest_info {
  uint64_t n_points, sdim, n_trees, max_leaf_size, height;  /* height: the tallest tree */
  int32_t scalar, device;
  double build_ms;          /* device time of reflections + builds (CUDA events) */
  uint64_t device_bytes;
} pico_b200_forest_info;

int pico_b200_forest_create(const void* pts, size_t n, size_t sdim, size_t stride_elems, int scalar,
                            size_t max_leaf_size, const void* rotations, size_t forest_size, int device,
                            pico_b200_forest** out);
void pico_b200_forest_destroy(pico_b200_forest* forest);
int pico_b200_forest_info_get(const pico_b200_forest* forest, pico_b200_forest_info* info);
/* the reflection vectors in use: forest_size * sdim scalars (kd_forest keeps them in rkd_tree_hh_data::rotation) */
int pico_b200_forest_rotations(const pico_b200_forest* forest, void* rotations_out);
/* tree `i` of the forest, borrowed (valid until the forest is destroyed): export / info of the per-tree structure */
int pico_b200_forest_tree(const pico_b200_forest* forest, size_t i, const pico_b200_tree** tree);
/*
 * kd_forest::search_nn (k = 1) / search_nearest with a search_knn visitor, one call per batch
 * (kd_forest.hpp:76-120). neighbors_out: nq * k records {int32 index, scalar distance}, ascending;
 * slots never filled keep index -1. flags: PICO_B200_DEVICE_POINTERS.
 */
int pico_b200_forest_knn(const pico_b200_forest* forest, const void* queries, size_t nq, size_t stride_elems, size_t k,
                         size_t max_leaves_visited, void* neighbors_out, unsigned flags, pico_b200_search_stats* stats);

/*
 * How the tree currently orders the queries of a batch before the thread-per-query kernels run (ordering never
 * changes a result, only memory coherence): 0 = not measured yet, 1 = batches arrived locally coherent (a scan, a
 * raster): every tile of 2048 consecutive queries is Z-ordered on its own, 2 = batches arrived in no useful order:
 * the whole batch is Z-ordered with a device-wide radix sort. Decided from the last measured batch.
 */
int pico_b200_tree_order_state(const pico_b200_tree* tree, int* state);

/*
 * Caller-provided CUDA stream (cudaStream_t) for the calling host thread; NULL restores the
 * default (a private stream per call). With a caller stream the searches are ordered on that
 * stream, so they compose with the caller's own kernels, events and CUDA graphs. NULL never means
 * "the default stream": pass cudaStreamLegacy ((cudaStream_t)0x1) or cudaStreamPerThread (0x2) to
 * order the searches after work pending on those.
 */
int pico_b200_set_stream(void* cuda_stream);

/*
 * Device-side profiling of the traversal kernels launched by the calling thread: between
 * begin and end every knn / radius traversal launch is bracketed by CUDA events on its
 * stream; end synchronises and returns their summed duration and their number.
 */
int pico_b200_profile_begin(void);
int pico_b200_profile_end(double* traversal_ms, uint64_t* traversal_launches);

/*
 * Measurement only (SURVEY.md §8d "leaf-scan HBM GB/s"): the leaf scan of kd_tree_search.hpp:54-59 in
 * isolation. Pass 1 walks every query (device pointer, Z-ordered like pico_b200_knn) from the root to its
 * first leaf (kd_tree_search.hpp:60-88) and stores the leaf's point range; pass 2 only streams those
 * contiguous point records through the search_nn visitor (search_visitor.hpp:41-65) and writes one
 * neighbour per query to `d_neighbors_out` (device pointer) — the nearest point inside the query's own
 * leaf. Both passes are timed with CUDA events over `repeats` runs; `scan_bytes` is what one pass-2
 * launch reads and writes (queries, ranges, point records, results). sdim <= 3, metric_l2_squared.
 */
int pico_b200_profile_leaf_scan(const pico_b200_tree* tree, const void* d_queries, size_t nq, size_t stride,
                                void* d_neighbors_out, int repeats, double* descend_ms, double* scan_ms,
                                uint64_t* scan_bytes);

void pico_b200_free(void* p);
/* frees a device buffer returned by pico_b200_radius / pico_b200_box under PICO_B200_DEVICE_POINTERS */
void pico_b200_free_device(void* p);

#ifdef __cplusplus
}
#endif
#endif /* PICO_B200_H_ */
