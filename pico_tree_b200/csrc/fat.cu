// fat.cu — the "search image" of a packed (sdim <= 3) tree: a second node array in which every
// subtree of at most `fat_limit` points is collapsed into ONE leaf record.
//
// Why: the reference's max_leaf_size (10 in every BASELINE config) is tuned for a CPU core. On the
// GPU the last two or three tree levels are where the thread-per-query traversal diverges: most far
// children that survive `visitor.max() >= dist` (kd_tree_search.hpp:99-103) are siblings of the
// first leaf, each costing a pop, a short descent and a 2-7 point scan at a handful of active lanes.
// Leaf order makes every subtree one contiguous run of pts4, so a collapsed subtree
// is scanned as one coalesced, warp-coherent run of float4 records instead.
//
// The node numbering is unchanged (pre-order indices of `nodes`), so a traversal may switch between
// the two arrays at any node. The exact single-neighbour search uses it (search.cu: nn_fat_kernel);
// the tree that is exported, saved and compared with the reference stays the real one.
#include "common.cuh"

namespace pico {
namespace {

template <typename T>
struct NodeIo;
template <>
struct NodeIo<float> {
  using Node = pico_b200_node_f32;
  __device__ static bool leaf(const Node& n) { return n.split_dim == PICO_B200_LEAF; }
  __device__ static int begin(const Node& n) { return n.a.begin_idx; }
  __device__ static int end(const Node& n) { return n.b.end_idx; }
  __device__ static void make_leaf(Node& n, int b, int e) {
    n.a.begin_idx = b;
    n.b.end_idx = e;
    n.right = PICO_B200_LEAF;
    n.split_dim = PICO_B200_LEAF;
  }
};
template <>
struct NodeIo<double> {
  using Node = pico_b200_node_f64;
  __device__ static bool leaf(const Node& n) { return n.split_dim == PICO_B200_LEAF; }
  __device__ static int begin(const Node& n) { return (int)n.a.begin_idx; }
  __device__ static int end(const Node& n) { return (int)n.b.end_idx; }
  __device__ static void make_leaf(Node& n, int b, int e) {
    n.a.begin_idx = b;
    n.b.end_idx = e;
    n.right = PICO_B200_LEAF;
    n.split_dim = PICO_B200_LEAF;
    n.pad_ = 0;
  }
};

// One thread per node. The points below branch i are [begin of its leftmost leaf, end of its
// rightmost leaf): two short walks (left child = i + 1, right child = `right`).
template <typename T>
__global__ void fat_nodes_kernel(const typename NodeIo<T>::Node* __restrict__ nodes, uint32_t n_nodes, int limit,
                                 typename NodeIo<T>::Node* __restrict__ fat, unsigned int* __restrict__ n_fat) {
  using Io = NodeIo<T>;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_nodes) return;
  typename Io::Node nd = nodes[i];
  if (!Io::leaf(nd)) {
    uint32_t l = i + 1, r = nd.right;
    typename Io::Node ln = nodes[l], rn = nodes[r];
    while (!Io::leaf(ln)) ln = nodes[++l];
    while (!Io::leaf(rn)) rn = nodes[r = rn.right];
    const int b = Io::begin(ln), e = Io::end(rn);
    if (e - b <= limit) {
      Io::make_leaf(nd, b, e);
      atomicAdd(n_fat, 1u);
    }
  }
  fat[i] = nd;
}

}  // namespace

// PICO_B200_FAT_LEAF: tuning hook (0 disables the search image; default kFatLeafPoints)
int fat_leaf_limit() {
  static const int v = [] {
    const char* e = getenv("PICO_B200_FAT_LEAF");
    const int x = e ? atoi(e) : -1;
    return (x >= 0 && x <= 4096) ? x : kFatLeafPoints;
  }();
  return v;
}

int build_fat_nodes(pico_b200_tree* t, cudaStream_t st) {
  cudaFree(t->d_fat_nodes);
  t->d_fat_nodes = nullptr;
  t->fat_limit = 0;
  const int limit = fat_leaf_limit();
  if (!t->packed() || t->n_nodes == 0 || limit <= 0 || (size_t)limit <= t->max_leaf_points) return 0;
  PICO_CUDA(cudaMalloc(&t->d_fat_nodes, t->n_nodes * t->node_size()));
  unsigned int* d_count = nullptr;
  PICO_CUDA(cudaMalloc(reinterpret_cast<void**>(&d_count), sizeof(unsigned int)));
  PICO_CUDA(cudaMemsetAsync(d_count, 0, sizeof(unsigned int), st));
  const unsigned blocks = (unsigned)((t->n_nodes + 255) / 256);
  if (t->scalar == PICO_B200_F32)
    fat_nodes_kernel<float><<<blocks, 256, 0, st>>>(static_cast<const pico_b200_node_f32*>(t->d_nodes),
                                                    (uint32_t)t->n_nodes, limit,
                                                    static_cast<pico_b200_node_f32*>(t->d_fat_nodes), d_count);
  else
    fat_nodes_kernel<double><<<blocks, 256, 0, st>>>(static_cast<const pico_b200_node_f64*>(t->d_nodes),
                                                     (uint32_t)t->n_nodes, limit,
                                                     static_cast<pico_b200_node_f64*>(t->d_fat_nodes), d_count);
  PICO_CUDA(cudaGetLastError());
  unsigned int n_fat = 0;
  PICO_CUDA(cudaMemcpyAsync(&n_fat, d_count, sizeof(n_fat), cudaMemcpyDeviceToHost, st));
  PICO_CUDA(cudaStreamSynchronize(st));
  cudaFree(d_count);
  if (n_fat == 0) {  // nothing to collapse (every leaf already holds more than `limit` points)
    cudaFree(t->d_fat_nodes);
    t->d_fat_nodes = nullptr;
    return 0;
  }
  t->fat_limit = limit;
  t->device_bytes += t->n_nodes * t->node_size();
  return 0;
}

}  // namespace pico
