// api.cu — the extern "C" surface declared in include/pico_b200.h.
#include <dlfcn.h>

#include <cstring>
#include <new>
#include <vector>

#include "common.cuh"

namespace pico {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

namespace {

int check_device(int device, int* sm_count) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    return fail(PICO_B200_ERR_NO_DEVICE,
                "no CUDA device available (libpico_b200 has no CPU fallback): " +
                    std::string(e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e)));
  }
  if (device < 0 || device >= count) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "device ordinal out of range");
  PICO_CUDA(cudaSetDevice(device));
  int major = 0, minor = 0, sms = 0;  // (cudaGetDeviceProperties costs milliseconds; three attributes do not)
  PICO_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  PICO_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
  PICO_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  if (major < 10)
    return fail(PICO_B200_ERR_NO_DEVICE, "libpico_b200 is built for sm_100a only; found sm_" + std::to_string(major) +
                                             std::to_string(minor));
  *sm_count = sms;
  // Per-call workspaces come from the stream-ordered allocator; keep freed blocks cached in
  // the pool instead of returning them to the driver at every synchronisation.
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    uint64_t keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  cudaGetLastError();
  return 0;
}

int check_common(const void* pts, size_t n, size_t sdim, size_t stride, int scalar, int metric) {
  if (!pts) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "points pointer is null");
  if (n == 0) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "empty point set (the reference asserts size() > 0)");
  if (sdim == 0) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "sdim must be > 0");
  if (stride < sdim) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "stride smaller than sdim");
  if (scalar != PICO_B200_F32 && scalar != PICO_B200_F64)
    return fail(PICO_B200_ERR_INVALID_ARGUMENT, "unknown scalar type");
  if (metric < PICO_B200_METRIC_L1 || metric > PICO_B200_METRIC_CUSTOM_EUCLIDEAN)
    return fail(PICO_B200_ERR_INVALID_ARGUMENT, "unknown metric");
  // metric_so2 reads coordinate 0, metric_se2_squared coordinates 0..2 (metric.hpp:203-208,229-238)
  if (metric == PICO_B200_METRIC_SO2 && sdim != 1)
    return fail(PICO_B200_ERR_INVALID_ARGUMENT, "metric_so2 needs sdim == 1");
  if (metric == PICO_B200_METRIC_SE2_SQUARED && sdim != 3)
    return fail(PICO_B200_ERR_INVALID_ARGUMENT, "metric_se2_squared needs sdim == 3");
  if (sdim > 0x7fff) return fail(PICO_B200_ERR_UNSUPPORTED, "sdim > 32767");
  return 0;
}

void release(pico_b200_tree* t) {
  if (!t) return;
  cudaSetDevice(t->device);
  cudaFree(t->d_nodes);
  cudaFree(t->d_pts);
  cudaFree(t->d_indices);
  cudaFree(t->d_root_box);
  cudaFree(t->d_outer);
  cudaFree(t->d_spans);
  cudaFree(t->d_fat_nodes);
  cudaFree(t->order_hint.d_stat);
  if (t->order_hint.h_stat) cudaFreeHost(t->order_hint.h_stat);
  delete t;
}

// Serialised image: header + root box + nodes + indices + points (device layout).
struct ImageHeader {
  uint64_t magic;  // "PICOB200"
  uint32_t version, scalar, metric, max_leaf_points;
  uint64_t n, sdim, n_nodes, n_leaves, height;
  double root_box_host[8];
};
constexpr uint64_t kMagic = 0x3030324a4f434950ull;

size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

// Structure check of an uploaded image (pico_b200_tree_deserialize, the receiving side of a broadcast): every
// branch links forward inside the array (pre-order: left = i + 1 < right < n_nodes), split dimensions exist, leaf
// ranges lie inside [0, n] and are ordered, indices are point numbers. flags[0] counts violations.
template <typename NodeT>
__global__ void validate_nodes_kernel(const NodeT* nodes, uint32_t n_nodes, uint32_t n, uint32_t sdim, uint32_t* bad) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_nodes) return;
  const NodeT nd = nodes[i];
  bool ok;
  if (nd.split_dim == PICO_B200_LEAF) {
    const long long b = (long long)nd.a.begin_idx, e = (long long)nd.b.end_idx;
    ok = nd.right == PICO_B200_LEAF && b >= 0 && b <= e && e <= (long long)n;
  } else {
    ok = nd.split_dim < sdim && nd.right > i + 1 && nd.right < n_nodes && i + 1 < n_nodes;
  }
  if (!ok) atomicAdd(bad, 1u);
}
__global__ void validate_indices_kernel(const int32_t* indices, uint32_t n, uint32_t* bad) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && (indices[i] < 0 || (uint32_t)indices[i] >= n)) atomicAdd(bad, 1u);
}
int validate_structure(const pico_b200_tree* t) {
  uint32_t* d_bad = nullptr;
  PICO_CUDA(cudaMalloc(reinterpret_cast<void**>(&d_bad), sizeof(uint32_t)));
  cudaMemset(d_bad, 0, sizeof(uint32_t));
  const unsigned nb = (unsigned)((t->n_nodes + 255) / 256);
  if (t->scalar == PICO_B200_F32)
    validate_nodes_kernel<<<nb, 256>>>(static_cast<const pico_b200_node_f32*>(t->d_nodes), (uint32_t)t->n_nodes,
                                       (uint32_t)t->n, (uint32_t)t->sdim, d_bad);
  else
    validate_nodes_kernel<<<nb, 256>>>(static_cast<const pico_b200_node_f64*>(t->d_nodes), (uint32_t)t->n_nodes,
                                       (uint32_t)t->n, (uint32_t)t->sdim, d_bad);
  validate_indices_kernel<<<(unsigned)((t->n + 255) / 256), 256>>>(t->d_indices, (uint32_t)t->n, d_bad);
  uint32_t bad = 0;
  const cudaError_t e = cudaMemcpy(&bad, d_bad, sizeof(bad), cudaMemcpyDeviceToHost);
  cudaFree(d_bad);
  if (e != cudaSuccess) return fail(PICO_B200_ERR_CUDA, std::string("tree image validation: ") + cudaGetErrorString(e));
  if (bad) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "tree image: corrupt node links, leaf ranges or indices");
  return 0;
}

size_t image_bytes(const pico_b200_tree* t) {
  return align16(sizeof(ImageHeader)) + align16(2 * t->sdim * t->scalar_size()) + align16(t->n_nodes * t->node_size()) +
         align16(t->n * 4) + align16(t->pts_bytes()) + align16(t->outer_bytes()) + align16(t->spans_bytes());
}

}  // namespace
}  // namespace pico

using namespace pico;

extern "C" {

const char* pico_b200_last_error(void) { return g_last_error.c_str(); }

int pico_b200_abi_version(void) { return PICO_B200_ABI_VERSION; }

int pico_b200_device_count(int* count) {
  if (!count) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "count is null");
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  if (e != cudaSuccess) {
    cudaGetLastError();
    c = 0;
  }
  *count = c;
  return 0;
}

int pico_b200_tree_create(const void* pts, size_t n, size_t sdim, size_t stride, int scalar, int metric, int rule,
                          int stop_kind, size_t stop_value, const void* bounds_min, const void* bounds_max, int device,
                          pico_b200_tree** out) {
  if (!out) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "out is null");
  *out = nullptr;
  PICO_TRY(check_common(pts, n, sdim, stride, scalar, metric));
  if (rule < 0 || rule > PICO_B200_RULE_MEDIAN_MAX_SIDE) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "unknown rule");
  if (stop_kind != PICO_B200_STOP_MAX_LEAF_SIZE && stop_kind != PICO_B200_STOP_MAX_LEAF_DEPTH)
    return fail(PICO_B200_ERR_INVALID_ARGUMENT, "unknown stop condition");
  if (stop_kind == PICO_B200_STOP_MAX_LEAF_SIZE && stop_value == 0)
    return fail(PICO_B200_ERR_INVALID_ARGUMENT, "max_leaf_size must be > 0 (kd_tree_builder.hpp:93)");
  if ((bounds_min == nullptr) != (bounds_max == nullptr))
    return fail(PICO_B200_ERR_INVALID_ARGUMENT, "bounds_min and bounds_max must both be given or both be null");
  int sms = 0;
  PICO_TRY(check_device(device, &sms));
  pico_b200_tree* t = new (std::nothrow) pico_b200_tree();
  if (!t) return fail(PICO_B200_ERR_OUT_OF_MEMORY, "host allocation failed");
  t->device = device;
  t->scalar = scalar;
  t->metric = metric;
  t->n = n;
  t->sdim = sdim;
  t->sm_count = sms;
  int rc;
  if (scalar == PICO_B200_F32)
    rc = build_tree<float>(t, static_cast<const float*>(pts), stride, rule, stop_kind, stop_value,
                           static_cast<const float*>(bounds_min), static_cast<const float*>(bounds_max));
  else
    rc = build_tree<double>(t, static_cast<const double*>(pts), stride, rule, stop_kind, stop_value,
                            static_cast<const double*>(bounds_min), static_cast<const double*>(bounds_max));
  if (rc) {
    release(t);
    return rc;
  }
  *out = t;
  return 0;
}

int pico_b200_tree_create_from_nodes(const void* pts, size_t n, size_t sdim, size_t stride, int scalar, int metric,
                                     const void* nodes, size_t n_nodes, const int32_t* indices, const void* root_box,
                                     const void* outer_bounds, int device, pico_b200_tree** out) {
  if (!out) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "out is null");
  *out = nullptr;
  PICO_TRY(check_common(pts, n, sdim, stride, scalar, metric));
  if (!nodes || !indices || !root_box) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "nodes/indices/root_box is null");
  int sms = 0;
  PICO_TRY(check_device(device, &sms));
  pico_b200_tree* t = new (std::nothrow) pico_b200_tree();
  if (!t) return fail(PICO_B200_ERR_OUT_OF_MEMORY, "host allocation failed");
  t->device = device;
  t->scalar = scalar;
  t->metric = metric;
  t->n = n;
  t->sdim = sdim;
  t->sm_count = sms;
  int rc;
  if (scalar == PICO_B200_F32)
    rc = upload_tree<float>(t, static_cast<const float*>(pts), stride, nodes, n_nodes, indices,
                            static_cast<const float*>(root_box), static_cast<const float*>(outer_bounds));
  else
    rc = upload_tree<double>(t, static_cast<const double*>(pts), stride, nodes, n_nodes, indices,
                             static_cast<const double*>(root_box), static_cast<const double*>(outer_bounds));
  if (rc) {
    release(t);
    return rc;
  }
  *out = t;
  return 0;
}

void pico_b200_tree_destroy(pico_b200_tree* tree) { release(tree); }

int pico_b200_tree_info_get(const pico_b200_tree* t, pico_b200_tree_info* info) {
  if (!t || !info) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "null argument");
  info->n_points = t->n;
  info->sdim = t->sdim;
  info->n_nodes = t->n_nodes;
  info->n_leaves = t->n_leaves;
  info->height = t->height;
  info->scalar = t->scalar;
  info->metric = t->metric;
  info->device = t->device;
  info->reserved_ = 0;
  info->build_ms = t->build_ms;
  info->device_bytes = t->device_bytes;
  return 0;
}

int pico_b200_tree_export_outer_bounds(const pico_b200_tree* t, void* outer_out) {
  if (!t || !outer_out) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "null argument");
  if (!t->outer_bytes()) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "only trees of a topological metric (and the trees of a kd_forest) keep outer bounds");
  PICO_CUDA(cudaSetDevice(t->device));
  PICO_CUDA(cudaMemcpy(outer_out, t->d_outer, t->outer_bytes(), cudaMemcpyDeviceToHost));
  return 0;
}

int pico_b200_tree_export(const pico_b200_tree* t, void* nodes_out, int32_t* indices_out, void* root_box_out) {
  if (!t) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "tree is null");
  PICO_CUDA(cudaSetDevice(t->device));
  if (nodes_out) PICO_CUDA(cudaMemcpy(nodes_out, t->d_nodes, t->n_nodes * t->node_size(), cudaMemcpyDeviceToHost));
  if (indices_out) PICO_CUDA(cudaMemcpy(indices_out, t->d_indices, t->n * 4, cudaMemcpyDeviceToHost));
  if (root_box_out)
    PICO_CUDA(cudaMemcpy(root_box_out, t->d_root_box, 2 * t->sdim * t->scalar_size(), cudaMemcpyDeviceToHost));
  return 0;
}

int pico_b200_knn(const pico_b200_tree* t, const void* queries, size_t nq, size_t stride, size_t k, double e,
                  void* neighbors_out, unsigned flags, pico_b200_search_stats* stats) {
  if (!t) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "tree is null");
  if (stats) memset(stats, 0, sizeof(*stats));
  if (nq == 0 || k == 0) return 0;
  if (!queries || !neighbors_out) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "null query or output pointer");
  if (stride < t->sdim) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "stride smaller than sdim");
  if (t->custom_metric())
    return fail(PICO_B200_ERR_UNSUPPORTED, "user-defined metric: nearest searches run in the host header, not here");
  if (t->scalar == PICO_B200_F32)
    return knn_batch<float>(t, static_cast<const float*>(queries), nq, stride, k, e,
                            static_cast<Neighbor<float>*>(neighbors_out), flags, stats);
  return knn_batch<double>(t, static_cast<const double*>(queries), nq, stride, k, e,
                           static_cast<Neighbor<double>*>(neighbors_out), flags, stats);
}

int pico_b200_radius(const pico_b200_tree* t, const void* queries, size_t nq, size_t stride, double radius, double e,
                     uint64_t* offsets_out, void** neighbors_out, unsigned flags, pico_b200_search_stats* stats) {
  if (!t || !offsets_out || !neighbors_out) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "null argument");
  if (stats) memset(stats, 0, sizeof(*stats));
  if (nq && !queries) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "null query pointer");
  if (stride < t->sdim) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "stride smaller than sdim");
  if (t->custom_metric())
    return fail(PICO_B200_ERR_UNSUPPORTED, "user-defined metric: nearest searches run in the host header, not here");
  if (t->scalar == PICO_B200_F32)
    return radius_batch<float>(t, static_cast<const float*>(queries), nq, stride, radius, e, offsets_out,
                               neighbors_out, flags, stats);
  return radius_batch<double>(t, static_cast<const double*>(queries), nq, stride, radius, e, offsets_out,
                              neighbors_out, flags, stats);
}

int pico_b200_box(const pico_b200_tree* t, const void* mins, const void* maxs, size_t nb, size_t stride,
                  uint64_t* offsets_out, int32_t** indices_out, unsigned flags, pico_b200_search_stats* stats) {
  if (!t || !offsets_out || !indices_out) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "null argument");
  if (stats) memset(stats, 0, sizeof(*stats));
  if (nb && (!mins || !maxs)) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "null box pointer");
  if (stride < t->sdim) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "stride smaller than sdim");
  if (t->metric == PICO_B200_METRIC_CUSTOM_TOPOLOGICAL)
    return fail(PICO_B200_ERR_UNSUPPORTED, "user-defined topological metric: which dimensions wrap is only known to the caller");
  if (t->scalar == PICO_B200_F32)
    return box_batch<float>(t, static_cast<const float*>(mins), static_cast<const float*>(maxs), nb, stride,
                            offsets_out, indices_out, flags, stats);
  return box_batch<double>(t, static_cast<const double*>(mins), static_cast<const double*>(maxs), nb, stride,
                           offsets_out, indices_out, flags, stats);
}

void pico_b200_free(void* p) { pico::release_result(p); }
void pico_b200_free_device(void* p) {
  // The buffers come from the stream-ordered pool and go back into it (a later call re-uses the block instead of
  // paying a multi-gigabyte cudaMalloc). Like cudaFree, this waits for everything the device was given so far.
  if (!p) return;
  cudaDeviceSynchronize();
  cudaFreeAsync(p, cudaStreamLegacy);
}

int pico_b200_tree_order_state(const pico_b200_tree* t, int* state) {
  if (!t || !state) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "null argument");
  *state = order_state(t);
  return 0;
}

int pico_b200_set_stream(void* cuda_stream) { return set_thread_stream(cuda_stream, cuda_stream != nullptr); }
int pico_b200_profile_begin(void) { return profile_begin(); }
int pico_b200_profile_end(double* traversal_ms, uint64_t* traversal_launches) {
  return profile_end(traversal_ms, traversal_launches);
}

int pico_b200_profile_leaf_scan(const pico_b200_tree* t, const void* d_queries, size_t nq, size_t stride,
                                void* d_neighbors_out, int repeats, double* descend_ms, double* scan_ms,
                                uint64_t* scan_bytes) {
  if (!t || !d_queries || !d_neighbors_out) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "null argument");
  if (stride < t->sdim) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "stride smaller than sdim");
  if (t->scalar == PICO_B200_F32)
    return leaf_scan_profile<float>(t, static_cast<const float*>(d_queries), nq, stride,
                                    static_cast<Neighbor<float>*>(d_neighbors_out), repeats, descend_ms, scan_ms,
                                    scan_bytes);
  return leaf_scan_profile<double>(t, static_cast<const double*>(d_queries), nq, stride,
                                   static_cast<Neighbor<double>*>(d_neighbors_out), repeats, descend_ms, scan_ms,
                                   scan_bytes);
}

// ---------------------------------------------------------------- (de)serialisation
int pico_b200_tree_serialize_size(const pico_b200_tree* t, uint64_t* bytes) {
  if (!t || !bytes) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "null argument");
  *bytes = image_bytes(t);
  return 0;
}

int pico_b200_tree_serialize(const pico_b200_tree* t, void* dst, int dst_is_device) {
  if (!t || !dst) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "null argument");
  PICO_CUDA(cudaSetDevice(t->device));
  ImageHeader h;
  memset(&h, 0, sizeof(h));
  h.magic = kMagic;
  h.version = PICO_B200_ABI_VERSION;
  h.scalar = (uint32_t)t->scalar;
  h.metric = (uint32_t)t->metric;
  h.max_leaf_points = (uint32_t)t->max_leaf_points;
  h.n = t->n;
  h.sdim = t->sdim;
  h.n_nodes = t->n_nodes;
  h.n_leaves = t->n_leaves;
  h.height = t->height;
  memcpy(h.root_box_host, t->root_box_host, sizeof(h.root_box_host));
  unsigned char* p = static_cast<unsigned char*>(dst);
  const cudaMemcpyKind kh = dst_is_device ? cudaMemcpyHostToDevice : cudaMemcpyHostToHost;
  const cudaMemcpyKind kd = dst_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  PICO_CUDA(cudaMemcpy(p, &h, sizeof(h), kh));
  p += align16(sizeof(ImageHeader));
  PICO_CUDA(cudaMemcpy(p, t->d_root_box, 2 * t->sdim * t->scalar_size(), kd));
  p += align16(2 * t->sdim * t->scalar_size());
  PICO_CUDA(cudaMemcpy(p, t->d_nodes, t->n_nodes * t->node_size(), kd));
  p += align16(t->n_nodes * t->node_size());
  PICO_CUDA(cudaMemcpy(p, t->d_indices, t->n * 4, kd));
  p += align16(t->n * 4);
  PICO_CUDA(cudaMemcpy(p, t->d_pts, t->pts_bytes(), kd));
  p += align16(t->pts_bytes());
  if (t->outer_bytes()) PICO_CUDA(cudaMemcpy(p, t->d_outer, t->outer_bytes(), kd));
  p += align16(t->outer_bytes());
  if (!t->packed()) PICO_CUDA(cudaMemcpy(p, t->d_spans, t->spans_bytes(), kd));
  return 0;
}

int pico_b200_tree_deserialize(const void* src, uint64_t bytes, int src_is_device, int device, pico_b200_tree** out) {
  if (!src || !out) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "null argument");
  *out = nullptr;
  int sms = 0;
  PICO_TRY(check_device(device, &sms));
  if (bytes < sizeof(ImageHeader)) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "image too small");
  ImageHeader h;
  PICO_CUDA(cudaMemcpy(&h, src, sizeof(h), src_is_device ? cudaMemcpyDeviceToHost : cudaMemcpyHostToHost));
  if (h.magic != kMagic || h.version != PICO_B200_ABI_VERSION)
    return fail(PICO_B200_ERR_INVALID_ARGUMENT, "not a pico_b200 tree image (magic/version mismatch)");
  // the header is data from outside: the same range checks as pico_b200_tree_create, plus the shape of a binary tree
  if ((h.scalar != PICO_B200_F32 && h.scalar != PICO_B200_F64) || h.metric > (uint32_t)PICO_B200_METRIC_CUSTOM_EUCLIDEAN ||
      h.sdim == 0 || h.sdim > 0x7fff || h.n == 0 || h.n >= 0x7fffffffull ||
      (h.metric == PICO_B200_METRIC_SO2 && h.sdim != 1) || (h.metric == PICO_B200_METRIC_SE2_SQUARED && h.sdim != 3))
    return fail(PICO_B200_ERR_INVALID_ARGUMENT, "tree image: header fields out of range");
  if (h.n_nodes == 0 || h.n_nodes != 2 * h.n_leaves - 1 || h.n_nodes > 2 * h.n + 1 || h.height >= h.n_nodes + 1)
    return fail(PICO_B200_ERR_INVALID_ARGUMENT, "tree image: node / leaf counts are not those of a binary tree");
  pico_b200_tree* t = new (std::nothrow) pico_b200_tree();
  if (!t) return fail(PICO_B200_ERR_OUT_OF_MEMORY, "host allocation failed");
  t->device = device;
  t->scalar = (int)h.scalar;
  t->metric = (int)h.metric;
  t->max_leaf_points = h.max_leaf_points;
  t->n = h.n;
  t->sdim = h.sdim;
  t->n_nodes = h.n_nodes;
  t->n_leaves = h.n_leaves;
  t->height = h.height;
  t->sm_count = sms;
  memcpy(t->root_box_host, h.root_box_host, sizeof(h.root_box_host));
  if (image_bytes(t) != bytes) {
    delete t;
    return fail(PICO_B200_ERR_INVALID_ARGUMENT, "image size does not match its header");
  }
  const cudaMemcpyKind kd = src_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  const unsigned char* p = static_cast<const unsigned char*>(src) + align16(sizeof(ImageHeader));
  auto up = [&](void** dstp, size_t sz) -> int {
    PICO_CUDA(cudaMalloc(dstp, sz ? sz : 16));
    PICO_CUDA(cudaMemcpy(*dstp, p, sz, kd));
    p += align16(sz);
    return 0;
  };
  int rc = up(&t->d_root_box, 2 * t->sdim * t->scalar_size());
  if (!rc) rc = up(&t->d_nodes, t->n_nodes * t->node_size());
  if (!rc) rc = up(reinterpret_cast<void**>(&t->d_indices), t->n * 4);
  if (!rc) rc = up(&t->d_pts, t->pts_bytes());
  if (!rc && t->topological()) rc = up(&t->d_outer, t->outer_bytes());
  if (!rc && !t->topological()) p += align16(t->outer_bytes());
  if (!rc && !t->packed()) rc = up(reinterpret_cast<void**>(&t->d_spans), t->spans_bytes());
  if (rc) {
    release(t);
    return rc;
  }
  t->device_bytes =
      t->pts_bytes() + t->n_nodes * t->node_size() + t->n * 4 + 2 * t->sdim * t->scalar_size() + t->outer_bytes() +
      t->spans_bytes();
  // links, leaf ranges, split dimensions and indices are checked on the device before any search may follow them
  rc = validate_structure(t);
  if (rc) {
    release(t);
    return rc;
  }
  rc = build_fat_nodes(t, nullptr);
  if (rc) {
    release(t);
    return rc;
  }
  *out = t;
  return 0;
}

}  // extern "C"

// ---------------------------------------------------------------- reference stream format
namespace {

template <typename T>
struct RefBranch {  // kd_tree_branch_single<T>, internal/kd_tree_node.hpp:42-50
  int32_t split_dim;
  T left_max;
  T right_min;
};
static_assert(sizeof(RefBranch<float>) == 12 && sizeof(RefBranch<double>) == 24, "raw struct sizes of the reference");

template <typename T>
struct RefBranchDouble {  // kd_tree_branch_double<T>, internal/kd_tree_node.hpp:52-59 (topological metrics)
  int32_t split_dim;
  T left_min;
  T left_max;
  T right_min;
  T right_max;
};
static_assert(sizeof(RefBranchDouble<float>) == 20 && sizeof(RefBranchDouble<double>) == 40, "raw struct sizes");

template <typename T>
size_t ref_branch_bytes(bool topo) {
  return topo ? sizeof(RefBranchDouble<T>) : sizeof(RefBranch<T>);
}

template <typename T>
uint64_t ref_stream_bytes(const pico_b200_tree* t) {
  const uint64_t branches = t->n_nodes - t->n_leaves;
  return 8 + 8 + 4 * (uint64_t)t->n + 2 * t->sdim * sizeof(T) + t->n_leaves * (1 + 8) +
         branches * (1 + ref_branch_bytes<T>(t->topological()));
}

template <typename T>
int ref_stream_write(const pico_b200_tree* t, unsigned char* dst) {
  using NodeT = typename NodeOf<T>::type;
  std::vector<NodeT> nodes(t->n_nodes);
  std::vector<int32_t> idx(t->n);
  std::vector<T> box(2 * t->sdim);
  PICO_TRY(pico_b200_tree_export(t, nodes.data(), idx.data(), box.data()));
  const bool topo = t->topological();
  std::vector<T> outer(topo ? 2 * t->n_nodes : 0);
  if (topo) PICO_TRY(pico_b200_tree_export_outer_bounds(t, outer.data()));
  size_t node_index = 0;
  unsigned char* p = dst;
  auto put = [&](const void* src, size_t n) {
    memcpy(p, src, n);
    p += n;
  };
  const uint64_t sdim = t->sdim, n = t->n;
  put(&sdim, 8);
  put(&n, 8);
  put(idx.data(), 4 * n);
  put(box.data(), 2 * sdim * sizeof(T));
  for (const NodeT& nd : nodes) {  // already pre-order
    if (nd.split_dim == PICO_B200_LEAF) {
      const unsigned char flag = 1;
      const int32_t be[2] = {(int32_t)nd.a.begin_idx, (int32_t)nd.b.end_idx};
      put(&flag, 1);
      put(be, 8);
    } else if (topo) {
      const unsigned char flag = 0;
      RefBranchDouble<T> b;
      memset(&b, 0, sizeof(b));
      b.split_dim = (int32_t)nd.split_dim;
      b.left_min = outer[2 * node_index];
      b.left_max = nd.a.left_max;
      b.right_min = nd.b.right_min;
      b.right_max = outer[2 * node_index + 1];
      put(&flag, 1);
      put(&b, sizeof(b));
    } else {
      const unsigned char flag = 0;
      RefBranch<T> b;
      memset(&b, 0, sizeof(b));
      b.split_dim = (int32_t)nd.split_dim;
      b.left_max = nd.a.left_max;
      b.right_min = nd.b.right_min;
      put(&flag, 1);
      put(&b, sizeof(b));
    }
    ++node_index;
  }
  return 0;
}

template <typename T>
int ref_stream_read(const unsigned char* src, uint64_t bytes, size_t n, size_t sdim, bool topo,
                    std::vector<int32_t>& idx, std::vector<T>& box, std::vector<typename NodeOf<T>::type>& nodes,
                    std::vector<T>& outer, uint64_t* consumed) {
  using NodeT = typename NodeOf<T>::type;
  const unsigned char* p = src;
  const unsigned char* end = src + bytes;
  auto get = [&](void* dst, size_t cnt) -> bool {
    if ((uint64_t)(end - p) < cnt) return false;
    memcpy(dst, p, cnt);
    p += cnt;
    return true;
  };
  uint64_t f_sdim = 0, f_n = 0;
  if (!get(&f_sdim, 8) || !get(&f_n, 8)) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "tree stream truncated");
  if (f_sdim != sdim || f_n != n)
    return fail(PICO_B200_ERR_INVALID_ARGUMENT, "tree stream does not match the point set (sdim / size differ)");
  idx.resize(n);
  box.resize(2 * sdim);
  if (!get(idx.data(), 4 * n) || !get(box.data(), 2 * sdim * sizeof(T)))
    return fail(PICO_B200_ERR_INVALID_ARGUMENT, "tree stream truncated");
  // read_node recursion (kd_tree_data.hpp:89-107) with an explicit stack of branches whose
  // right child is still to come
  nodes.clear();
  outer.clear();
  std::vector<uint32_t> pending;
  bool done = false;
  while (!done) {
    unsigned char flag;
    if (!get(&flag, 1)) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "tree stream truncated");
    NodeT nd;
    memset(&nd, 0, sizeof(nd));
    if (flag) {
      int32_t be[2];
      if (!get(be, 8)) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "tree stream truncated");
      nd.a.begin_idx = be[0];
      nd.b.end_idx = be[1];
      nd.right = PICO_B200_LEAF;
      nd.split_dim = PICO_B200_LEAF;
      nodes.push_back(nd);
      if (topo) outer.insert(outer.end(), 2, T(0));
      // a finished subtree: the innermost pending branch gets its right child next
      if (pending.empty()) {
        done = true;
      } else {
        nodes[pending.back()].right = (uint32_t)nodes.size();
        pending.pop_back();
      }
    } else if (topo) {
      RefBranchDouble<T> b;
      if (!get(&b, sizeof(b))) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "tree stream truncated");
      nd.a.left_max = b.left_max;
      nd.b.right_min = b.right_min;
      nd.split_dim = (uint32_t)b.split_dim;
      nd.right = 0;
      pending.push_back((uint32_t)nodes.size());
      nodes.push_back(nd);
      outer.push_back(b.left_min);
      outer.push_back(b.right_max);
    } else {
      RefBranch<T> b;
      if (!get(&b, sizeof(b))) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "tree stream truncated");
      nd.a.left_max = b.left_max;
      nd.b.right_min = b.right_min;
      nd.split_dim = (uint32_t)b.split_dim;
      nd.right = 0;
      pending.push_back((uint32_t)nodes.size());
      nodes.push_back(nd);
    }
    if (nodes.size() > 0x7ffffff0u) return fail(PICO_B200_ERR_UNSUPPORTED, "too many nodes");
  }
  if (consumed) *consumed = (uint64_t)(p - src);
  return 0;
}

}  // namespace

extern "C" {

int pico_b200_tree_save_size(const pico_b200_tree* t, uint64_t* bytes) {
  if (!t || !bytes) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "null argument");
  *bytes = t->scalar == PICO_B200_F32 ? ref_stream_bytes<float>(t) : ref_stream_bytes<double>(t);
  return 0;
}

int pico_b200_tree_save(const pico_b200_tree* t, void* dst) {
  if (!t || !dst) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "null argument");
  return t->scalar == PICO_B200_F32 ? ref_stream_write<float>(t, static_cast<unsigned char*>(dst))
                                    : ref_stream_write<double>(t, static_cast<unsigned char*>(dst));
}

int pico_b200_tree_load(const void* pts, size_t n, size_t sdim, size_t stride, int scalar, int metric,
                        const void* stream, uint64_t stream_bytes, int device, pico_b200_tree** out,
                        uint64_t* consumed) {
  if (!out) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "out is null");
  *out = nullptr;
  if (!stream) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "stream is null");
  PICO_TRY(check_common(pts, n, sdim, stride, scalar, metric));
  const bool topo = metric >= PICO_B200_METRIC_SO2 && metric <= PICO_B200_METRIC_CUSTOM_TOPOLOGICAL;
  std::vector<int32_t> idx;
  if (scalar == PICO_B200_F32) {
    std::vector<float> box, outer;
    std::vector<pico_b200_node_f32> nodes;
    PICO_TRY(ref_stream_read<float>(static_cast<const unsigned char*>(stream), stream_bytes, n, sdim, topo, idx, box,
                                    nodes, outer, consumed));
    return pico_b200_tree_create_from_nodes(pts, n, sdim, stride, scalar, metric, nodes.data(), nodes.size(),
                                            idx.data(), box.data(), topo ? outer.data() : nullptr, device, out);
  }
  std::vector<double> box, outer;
  std::vector<pico_b200_node_f64> nodes;
  PICO_TRY(ref_stream_read<double>(static_cast<const unsigned char*>(stream), stream_bytes, n, sdim, topo, idx, box,
                                   nodes, outer, consumed));
  return pico_b200_tree_create_from_nodes(pts, n, sdim, stride, scalar, metric, nodes.data(), nodes.size(),
                                          idx.data(), box.data(), topo ? outer.data() : nullptr, device, out);
}

// ---------------------------------------------------------------- NCCL broadcast
// NCCL is resolved at run time (dlopen) so that the library has no link-time dependency on
// it; callers that already hold a communicator (ncclComm_t) pass it in.
int pico_b200_tree_broadcast(pico_b200_tree** tree, void* nccl_comm, int rank, int root, int device) {
  if (!tree || !nccl_comm) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "null argument");
  typedef int (*bcast_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
  static bcast_fn bcast = nullptr;
  if (!bcast) {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return fail(PICO_B200_ERR_NCCL, std::string("cannot load libnccl: ") + dlerror());
    bcast = reinterpret_cast<bcast_fn>(dlsym(h, "ncclBroadcast"));
    if (!bcast) return fail(PICO_B200_ERR_NCCL, "ncclBroadcast not found in libnccl");
  }
  int sms = 0;
  PICO_TRY(check_device(device, &sms));
  if (rank == root && !*tree) return fail(PICO_B200_ERR_INVALID_ARGUMENT, "root rank has no tree to broadcast");
  uint64_t bytes = 0;
  if (rank == root) bytes = image_bytes(*tree);
  uint64_t* d_size = nullptr;
  PICO_CUDA(cudaMalloc(&d_size, 8));
  PICO_CUDA(cudaMemcpy(d_size, &bytes, 8, cudaMemcpyHostToDevice));
  const int ncclChar = 0;
  if (bcast(d_size, d_size, 8, ncclChar, root, nccl_comm, 0) != 0) {
    cudaFree(d_size);
    return fail(PICO_B200_ERR_NCCL, "ncclBroadcast(size) failed");
  }
  PICO_CUDA(cudaMemcpy(&bytes, d_size, 8, cudaMemcpyDeviceToHost));
  cudaFree(d_size);
  void* image = nullptr;
  PICO_CUDA(cudaMalloc(&image, bytes));
  int rc = 0;
  if (rank == root) rc = pico_b200_tree_serialize(*tree, image, 1);
  if (!rc && bcast(image, image, bytes, ncclChar, root, nccl_comm, 0) != 0)
    rc = fail(PICO_B200_ERR_NCCL, "ncclBroadcast(image) failed");
  if (!rc) rc = cudaStreamSynchronize(0) == cudaSuccess ? 0 : fail(PICO_B200_ERR_CUDA, "sync after broadcast failed");
  if (!rc && rank != root) {
    if (*tree) {
      pico_b200_tree_destroy(*tree);
      *tree = nullptr;
    }
    rc = pico_b200_tree_deserialize(image, bytes, 1, device, tree);
  }
  cudaFree(image);
  return rc;
}

}  // extern "C"
