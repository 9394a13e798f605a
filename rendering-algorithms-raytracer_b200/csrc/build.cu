// build.cu — acceleration-structure build ON THE GPU: an LBVH over the scene's static triangles.
//
// Replaces, as an option, the host build the reference keeps (BVH::build, src/BVH.cpp:457-1106: binned SAH, seconds for a
// million triangles) when a scene description arrives with root == MIRO_GPU_ROOT_BUILD_ON_DEVICE: Morton codes of the
// triangle centroids -> radix sort (cub) -> binary radix tree (Karras 2012: every internal node covers a contiguous range of
// the sorted triangles) -> bounds bottom-up -> collapse to the 4-wide, 64-byte quantized device node of traverse.cuh,
// top-down, one kernel launch per level of the wide tree (a range of <= 4 triangles is a leaf, so no data is moved beyond the
// sort).  Closest-hit results do not depend on the tree; an LBVH costs more node visits per ray than the host's SAH tree and
// builds ~1000x faster.
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <float.h>
#include <stdlib.h>
#include "context.cuh"

namespace miro {

namespace {

__device__ __forceinline__ uint32_t expand_bits(uint32_t v) {      // 10 bits -> every third bit
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

__device__ __forceinline__ int float_order(float f) { const int i = __float_as_int(f); return i ^ ((i >> 31) & 0x7fffffff); }
__device__ __forceinline__ float order_float(int k) { return __int_as_float(k ^ ((k >> 31) & 0x7fffffff)); }

struct Bounds6 { int lo[3], hi[3]; };       // order-preserving int images of floats, for atomicMin / atomicMax

__global__ void k_scene_bounds(const float4* __restrict__ tris, uint32_t n, Bounds6* __restrict__ b) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    if (i < n) {
        const float4 p0 = tris[3 * (size_t)i], p1 = tris[3 * (size_t)i + 1], p2 = tris[3 * (size_t)i + 2];
        const float c[3] = {(fminf(p0.x, fminf(p1.x, p2.x)) + fmaxf(p0.x, fmaxf(p1.x, p2.x))) * 0.5f,
                            (fminf(p0.y, fminf(p1.y, p2.y)) + fmaxf(p0.y, fmaxf(p1.y, p2.y))) * 0.5f,
                            (fminf(p0.z, fminf(p1.z, p2.z)) + fmaxf(p0.z, fmaxf(p1.z, p2.z))) * 0.5f};
        for (int k = 0; k < 3; ++k) { lo[k] = c[k]; hi[k] = c[k]; }
    }
    for (int k = 0; k < 3; ++k) {
        for (int o = 16; o > 0; o >>= 1) { lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o)); hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o)); }
        if ((threadIdx.x & 31) == 0 && lo[k] <= hi[k]) { atomicMin(&b->lo[k], float_order(lo[k])); atomicMax(&b->hi[k], float_order(hi[k])); }
    }
}

__global__ void k_morton(const float4* __restrict__ tris, uint32_t n, const Bounds6* __restrict__ b, unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p0 = tris[3 * (size_t)i], p1 = tris[3 * (size_t)i + 1], p2 = tris[3 * (size_t)i + 2];
    const float c[3] = {(fminf(p0.x, fminf(p1.x, p2.x)) + fmaxf(p0.x, fmaxf(p1.x, p2.x))) * 0.5f,
                        (fminf(p0.y, fminf(p1.y, p2.y)) + fmaxf(p0.y, fmaxf(p1.y, p2.y))) * 0.5f,
                        (fminf(p0.z, fminf(p1.z, p2.z)) + fmaxf(p0.z, fmaxf(p1.z, p2.z))) * 0.5f};
    uint32_t q[3];
    for (int k = 0; k < 3; ++k) {
        const float lo = order_float(b->lo[k]), hi = order_float(b->hi[k]);
        const float e = hi - lo;
        const float u = e > 0.f ? (c[k] - lo) / e : 0.f;
        q[k] = (uint32_t)fminf(fmaxf(u * 1024.f, 0.f), 1023.f);
    }
    const uint32_t code = (expand_bits(q[0]) << 2) | (expand_bits(q[1]) << 1) | expand_bits(q[2]);
    keys[i] = ((unsigned long long)code << 32) | i;        // the index makes every key unique (Karras 2012, sec. 4)
    vals[i] = i;
}

// Karras 2012: internal node i of the binary radix tree over n sorted, unique keys.
__device__ __forceinline__ int delta(const unsigned long long* __restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    return __clzll(keys[i] ^ keys[j]);
}

struct BinTree {
    int* left;  int* right;     // children of internal node i: >= 0 internal node, < 0 leaf ~index (sorted position)
    int* parent;                // of internal nodes (root: -1) ...
    int* leaf_parent;           // ... and of leaves
    int* first; int* last;      // range of sorted triangles covered by internal node i
    float* lo;  float* hi;      // 3 floats per internal node
    int* visits;                // bottom-up arrival counters
    float* cost;                // SAH cost of the subtree as built (per internal node)
    int* is_leaf;               // the subtree is cheaper as ONE leaf of <= 4 triangles than split (decided bottom-up)
};

__global__ void k_radix_tree(const unsigned long long* __restrict__ keys, int n, BinTree t) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int s = lmax >> 1; s >= 1; s >>= 1) if (delta(keys, n, i, i + (l + s) * d) > dmin) l += s;
    const int j = i + l * d;
    const int dnode = delta(keys, n, i, j);
    int s = 0;
    for (int div = 2;; div <<= 1) {
        const int step = (l + div - 1) / div;
        if (delta(keys, n, i, i + (s + step) * d) > dnode) s += step;
        if (step <= 1) break;
    }
    const int gamma = i + s * d + min(d, 0);
    const int lo_i = min(i, j), hi_i = max(i, j);
    const int lc = (lo_i == gamma) ? ~gamma : gamma;
    const int rc = (hi_i == gamma + 1) ? ~(gamma + 1) : gamma + 1;
    t.left[i] = lc; t.right[i] = rc; t.first[i] = lo_i; t.last[i] = hi_i;
    if (lc >= 0) t.parent[lc] = i; else t.leaf_parent[~lc] = i;
    if (rc >= 0) t.parent[rc] = i; else t.leaf_parent[~rc] = i;
    if (i == 0) t.parent[0] = -1;
}

// Bounds of every internal node, bottom-up: the second thread to arrive at a node owns it (its two children are complete).
__global__ void k_fit(const float4* __restrict__ tris_sorted, int n, BinTree t, float level_cost) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int node = t.leaf_parent[i];
    while (node >= 0) {
        __threadfence();
        if (atomicAdd(&t.visits[node], 1) == 0) return;
        float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        const int c[2] = {t.left[node], t.right[node]};
        float child_term = 0.f;           // sum over the two children of area * subtree cost
        for (int k = 0; k < 2; ++k) {
            float clo[3], chi[3], ccost;
            if (c[k] < 0) {
                const float4* p = tris_sorted + TRI_F4 * (size_t)(~c[k]);
                clo[0] = fminf(p[0].x, fminf(p[1].x, p[2].x)); clo[1] = fminf(p[0].y, fminf(p[1].y, p[2].y)); clo[2] = fminf(p[0].z, fminf(p[1].z, p[2].z));
                chi[0] = fmaxf(p[0].x, fmaxf(p[1].x, p[2].x)); chi[1] = fmaxf(p[0].y, fmaxf(p[1].y, p[2].y)); chi[2] = fmaxf(p[0].z, fmaxf(p[1].z, p[2].z));
                ccost = 1.f;
            } else {
                // written by another thread before its atomicAdd: read through L2
                for (int a = 0; a < 3; ++a) { clo[a] = __ldcg(&t.lo[3 * c[k] + a]); chi[a] = __ldcg(&t.hi[3 * c[k] + a]); }
                ccost = __ldcg(&t.cost[c[k]]);
            }
            const float dx = chi[0] - clo[0], dy = chi[1] - clo[1], dz = chi[2] - clo[2];
            child_term += (dx * dy + dy * dz + dz * dx) * ccost;
            for (int a = 0; a < 3; ++a) { lo[a] = fminf(lo[a], clo[a]); hi[a] = fmaxf(hi[a], chi[a]); }
        }
        for (int a = 0; a < 3; ++a) { t.lo[3 * node + a] = lo[a]; t.hi[3 * node + a] = hi[a]; }
        // surface-area heuristic, bottom-up: keep the split only when it is cheaper than one leaf holding the whole range
        const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        const float area = fmaxf(dx * dy + dy * dz + dz * dx, 1e-30f);
        const int count = t.last[node] - t.first[node] + 1;
        const float split_cost = level_cost + child_term / area;    // level_cost: one binary level of a 4-wide node, in triangle tests
        const bool leaf = count <= (int)MIRO_GPU_MAX_LEAF && (float)count <= split_cost;
        t.cost[node] = leaf ? (float)count : split_cost;
        t.is_leaf[node] = leaf ? 1 : 0;
        node = t.parent[node];
    }
}

__global__ void k_gather_tris(const float4* __restrict__ in, const uint32_t* __restrict__ perm, uint32_t n, float4* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t s = (size_t)perm[i] * 3;
    float4* o = out + TRI_F4 * (size_t)i;       // device triangle records (traverse.cuh, TRI_F4)
    o[0] = in[s]; o[1] = in[s + 1]; o[2] = in[s + 2];
    if (TRI_F4 == 4) o[TRI_F4 - 1] = make_float4(0.f, 0.f, 0.f, 0.f);
}
__global__ void k_gather_prims(const miro_gpu_prim* __restrict__ in, const uint32_t* __restrict__ perm, uint32_t n, miro_gpu_prim* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = in[perm[i]];
}

// One level of the wide tree: queue entry = (binary node that becomes a wide node, index of that wide node).
struct WideItem { int bin; int wide; };

__device__ __forceinline__ int range_size(const BinTree& t, int c) { return c < 0 ? 1 : t.last[c] - t.first[c] + 1; }
__device__ __forceinline__ bool is_leaf_child(const BinTree& t, int c) { return c < 0 || t.is_leaf[c]; }

__global__ void k_collapse(BinTree t, const float4* __restrict__ tris_sorted, const WideItem* __restrict__ in, int n_in,
                           WideItem* __restrict__ out, int* __restrict__ n_out, int* __restrict__ n_wide, DeviceNode* __restrict__ nodes) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_in) return;
    const WideItem it = in[i];
    // up to four children: open the child covering the most triangles until four are held or none can be opened
    int kids[4]; int nk = 2;
    kids[0] = t.left[it.bin]; kids[1] = t.right[it.bin];
    while (nk < 4) {
        int best = -1, best_sz = 0;
        for (int k = 0; k < nk; ++k) { const int sz = range_size(t, kids[k]); if (!is_leaf_child(t, kids[k]) && sz > best_sz) { best_sz = sz; best = k; } }
        if (best < 0) break;
        const int c = kids[best];
        kids[best] = t.left[c]; kids[nk++] = t.right[c];
    }
    miro_gpu_node nd;
    for (int k = 0; k < 4; ++k) {
        nd.lo_x[k] = nd.lo_y[k] = nd.lo_z[k] = FLT_MAX; nd.hi_x[k] = nd.hi_y[k] = nd.hi_z[k] = -FLT_MAX;
        nd.child[k] = MIRO_GPU_CHILD_EMPTY; nd.reserved[k] = 0;
    }
    for (int k = 0; k < nk; ++k) {
        const int c = kids[k];
        float lo[3], hi[3];
        int first, count;
        if (c < 0) {
            first = ~c; count = 1;
            const float4* p = tris_sorted + TRI_F4 * (size_t)first;
            lo[0] = fminf(p[0].x, fminf(p[1].x, p[2].x)); lo[1] = fminf(p[0].y, fminf(p[1].y, p[2].y)); lo[2] = fminf(p[0].z, fminf(p[1].z, p[2].z));
            hi[0] = fmaxf(p[0].x, fmaxf(p[1].x, p[2].x)); hi[1] = fmaxf(p[0].y, fmaxf(p[1].y, p[2].y)); hi[2] = fmaxf(p[0].z, fmaxf(p[1].z, p[2].z));
        } else {
            first = t.first[c]; count = t.last[c] - first + 1;
            for (int a = 0; a < 3; ++a) { lo[a] = t.lo[3 * c + a]; hi[a] = t.hi[3 * c + a]; }
        }
        nd.lo_x[k] = lo[0]; nd.lo_y[k] = lo[1]; nd.lo_z[k] = lo[2]; nd.hi_x[k] = hi[0]; nd.hi_y[k] = hi[1]; nd.hi_z[k] = hi[2];
        if (is_leaf_child(t, c)) nd.child[k] = MIRO_GPU_LEAF(MIRO_GPU_KIND_TRI, first, count);
        else {
            const int w = atomicAdd(n_wide, 1);
            nd.child[k] = w;
            const int q = atomicAdd(n_out, 1);
            out[q].bin = c; out[q].wide = w;
        }
    }
    nodes[it.wide] = compress_node(nd);
}

template <class T>
struct Scratch {
    T* p = nullptr;
    cudaError_t alloc(size_t n) { return cudaMalloc((void**)&p, std::max<size_t>(n, 1) * sizeof(T)); }
    ~Scratch() { if (p) cudaFree(p); }
};

}  // namespace

// Builds the tree over d_tris_in (n triangles, caller's order).  Outputs (device memory, owned by the caller through
// ctx->scene_allocs): the node array, the triangles in leaf order, perm[i] = caller's index of sorted triangle i.
int build_lbvh_on_device(miro_gpu_ctx* ctx, const float4* d_tris_in, uint32_t n, const DeviceNode** out_nodes, uint32_t* out_n_nodes,
                         int32_t* out_root, const float4** out_tris, const uint32_t** out_perm) {
    cudaStream_t s = ctx->stream;
    auto keep = [&](void* p) { ctx->scene_allocs.push_back(p); };
    float4* tris_sorted = nullptr; uint32_t* perm = nullptr; DeviceNode* nodes = nullptr;
    MIRO_CUDA(ctx, cudaMalloc((void**)&tris_sorted, (size_t)std::max<uint32_t>(n, 1) * TRI_F4 * sizeof(float4))); keep(tris_sorted);
    MIRO_CUDA(ctx, cudaMalloc((void**)&perm, (size_t)std::max<uint32_t>(n, 1) * 4)); keep(perm);
    *out_tris = tris_sorted; *out_perm = perm;
    const int B = 256;
    auto grid = [&](size_t k) { return (unsigned)((k + B - 1) / B); };
    if (n <= MIRO_GPU_MAX_LEAF) {        // the whole scene is one leaf (cf. src/BVH.cpp:118-132)
        MIRO_CUDA(ctx, cudaMemsetAsync(tris_sorted, 0, (size_t)std::max<uint32_t>(n, 1) * TRI_F4 * sizeof(float4), s));
        if (n) MIRO_CUDA(ctx, cudaMemcpy2DAsync(tris_sorted, TRI_F4 * sizeof(float4), d_tris_in, 48, 48, n, cudaMemcpyDeviceToDevice, s));      // 48-byte ABI records -> device records
        std::vector<uint32_t> id(n); for (uint32_t i = 0; i < n; ++i) id[i] = i;
        MIRO_CUDA(ctx, cudaMemcpyAsync(perm, id.data(), (size_t)n * 4, cudaMemcpyHostToDevice, s));
        MIRO_CUDA(ctx, cudaStreamSynchronize(s));
        MIRO_CUDA(ctx, cudaMalloc((void**)&nodes, sizeof(DeviceNode))); keep(nodes);
        *out_nodes = nodes; *out_n_nodes = 0; *out_root = n ? MIRO_GPU_LEAF(MIRO_GPU_KIND_TRI, 0, n) : MIRO_GPU_CHILD_EMPTY;
        return MIRO_GPU_OK;
    }
    Scratch<Bounds6> bounds; Scratch<unsigned long long> keys_in, keys_out; Scratch<uint32_t> vals_in; Scratch<unsigned char> tmp;
    Scratch<int> ints; Scratch<float> boxes; Scratch<WideItem> queue; Scratch<int> counters;
    MIRO_CUDA(ctx, bounds.alloc(1)); MIRO_CUDA(ctx, keys_in.alloc(n)); MIRO_CUDA(ctx, keys_out.alloc(n)); MIRO_CUDA(ctx, vals_in.alloc(n));
    MIRO_CUDA(ctx, ints.alloc((size_t)n * 8)); MIRO_CUDA(ctx, boxes.alloc((size_t)n * 7)); MIRO_CUDA(ctx, queue.alloc((size_t)n * 2)); MIRO_CUDA(ctx, counters.alloc(4));
    // 1. centroid bounds, Morton keys
    const Bounds6 init = {{INT_MAX, INT_MAX, INT_MAX}, {INT_MIN, INT_MIN, INT_MIN}};
    MIRO_CUDA(ctx, cudaMemcpyAsync(bounds.p, &init, sizeof(init), cudaMemcpyHostToDevice, s));
    k_scene_bounds<<<grid(n), B, 0, s>>>(d_tris_in, n, bounds.p);
    k_morton<<<grid(n), B, 0, s>>>(d_tris_in, n, bounds.p, keys_in.p, vals_in.p);
    // 2. sort
    size_t tmp_bytes = 0;
    MIRO_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_in.p, keys_out.p, vals_in.p, perm, (int)n, 0, 64, s));
    MIRO_CUDA(ctx, tmp.alloc(tmp_bytes));
    MIRO_CUDA(ctx, cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, keys_in.p, keys_out.p, vals_in.p, perm, (int)n, 0, 64, s));
    k_gather_tris<<<grid(n), B, 0, s>>>(d_tris_in, perm, n, tris_sorted);
    // 3. binary radix tree + bounds
    BinTree t;
    t.left = ints.p; t.right = ints.p + n; t.parent = ints.p + 2 * (size_t)n; t.leaf_parent = ints.p + 3 * (size_t)n;
    t.first = ints.p + 4 * (size_t)n; t.last = ints.p + 5 * (size_t)n; t.visits = ints.p + 6 * (size_t)n;
    t.lo = boxes.p; t.hi = boxes.p + 3 * (size_t)n; t.cost = boxes.p + 6 * (size_t)n; t.is_leaf = ints.p + 7 * (size_t)n;
    MIRO_CUDA(ctx, cudaMemsetAsync(t.visits, 0, (size_t)n * sizeof(int), s));
    k_radix_tree<<<grid(n - 1), B, 0, s>>>(keys_out.p, (int)n, t);
    float level_cost = 1.0f;
    if (const char* e = getenv("MIRO_LBVH_LEVEL_COST")) { const float v = (float)atof(e); if (v > 0.f) level_cost = v; }      // tuning aid
    k_fit<<<grid(n), B, 0, s>>>(tris_sorted, (int)n, t, level_cost);
    // 4. collapse to the wide tree, one launch per level; a wide node stands for a binary internal node: at most n - 1 of them
    const size_t max_wide = (size_t)n;
    MIRO_CUDA(ctx, cudaMalloc((void**)&nodes, max_wide * sizeof(DeviceNode))); keep(nodes);
    WideItem* q[2] = {queue.p, queue.p + n};
    const WideItem root_item = {0, 0};
    MIRO_CUDA(ctx, cudaMemcpyAsync(q[0], &root_item, sizeof(root_item), cudaMemcpyHostToDevice, s));
    int h_counts[2] = {0, 1};      // [0] items queued for the next level, [1] wide nodes allocated
    MIRO_CUDA(ctx, cudaMemcpyAsync(counters.p, h_counts, sizeof(h_counts), cudaMemcpyHostToDevice, s));
    int n_in = 1, cur = 0, levels = 0;
    while (n_in > 0) {
        k_collapse<<<grid((size_t)n_in), B, 0, s>>>(t, tris_sorted, q[cur], n_in, q[cur ^ 1], counters.p, counters.p + 1, nodes);
        MIRO_CUDA(ctx, cudaMemcpyAsync(h_counts, counters.p, sizeof(h_counts), cudaMemcpyDeviceToHost, s));
        MIRO_CUDA(ctx, cudaMemsetAsync(counters.p, 0, sizeof(int), s));
        MIRO_CUDA(ctx, cudaStreamSynchronize(s));
        n_in = h_counts[0]; cur ^= 1;
        if (++levels > 128) return set_error(ctx, MIRO_GPU_ECUDA, "device BVH build did not terminate");
    }
    MIRO_CUDA(ctx, cudaGetLastError());
    ctx->launches += 5 + levels;
    *out_nodes = nodes; *out_n_nodes = (uint32_t)h_counts[1]; *out_root = 0;
    ctx->build_levels = levels;
    return MIRO_GPU_OK;
}

int reorder_prims_on_device(miro_gpu_ctx* ctx, const miro_gpu_prim* d_in, const uint32_t* d_perm, uint32_t n, const miro_gpu_prim** out) {
    miro_gpu_prim* p = nullptr;
    MIRO_CUDA(ctx, cudaMalloc((void**)&p, (size_t)std::max<uint32_t>(n, 1) * sizeof(miro_gpu_prim)));
    ctx->scene_allocs.push_back(p);
    if (n) k_gather_prims<<<(n + 255) / 256, 256, 0, ctx->stream>>>(d_in, d_perm, n, p);
    MIRO_CUDA(ctx, cudaGetLastError());
    *out = p;
    return MIRO_GPU_OK;
}

}  // namespace miro
