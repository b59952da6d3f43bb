// tray_lbvh.cuh -- SURVEY 8(f)-2: the closest-hit BVH of large scenes built on the device (LBVH: Morton order + the
// parallel radix-tree construction of Karras 2012), next to the host's median-split build.
//
// The tree only has to be CONSERVATIVE: traversal (bvh_closest_hit) runs the exact Sphere.Hit on every sphere it reaches
// and resolves ties by (t, id), so any tree over the same spheres gives the same bits as the reference's linear scan.
//
//   lbvh_keys_kernel     64-bit key per sphere: 48-bit Morton code of the centre (16 bits per axis) | 16-bit sphere id
//                        (unique keys: no duplicate handling needed in the radix tree)
//   lbvh_bitonic_kernel  bitonic sorting network over the padded key array, one compare-exchange stage per launch
//   lbvh_tree_kernel     one thread per internal node: its key range, split position, children
//   lbvh_boxes_kernel    one thread per sphere, bottom-up: leaf boxes (padded like the host build), unions at the second
//                        arrival, ranges of <= 4 spheres collapse into one leaf, near/far child order, subtree depth
#pragma once
#include "tray_device.cuh"

namespace tray {

struct LbvhArgs {
    int m;                              // spheres in the tree
    const int* ids;                     // their scene indices
    const double4* geo;                 // cx, cy, cz, r*r
    const double* radius;
    double lo[3], inv_extent[3];        // centre bounds -> [0,1)
    unsigned long long* keys;           // padded to a power of two
    int n_keys;
    int* parent;                        // per node (internal 0..m-2, single-sphere leaves m-1..2m-2)
    int2* range;                        // per internal node: first, last key index
    int2* child;                        // per internal node
    int* visits;                        // per internal node
    int* depth;                         // per node
    BvhNode* nodes;                     // 2m-1
    int* leaf_ids;                      // m, in key order
};

__device__ __forceinline__ unsigned long long lbvh_spread16(unsigned v) {  // ...abc -> ..a..b..c
    unsigned long long x = v & 0xFFFFu;
    x = (x | (x << 16)) & 0x0000FF0000FFULL;
    x = (x | (x << 8)) & 0x00F00F00F00FULL;
    x = (x | (x << 4)) & 0x0C30C30C30C3ULL;
    x = (x | (x << 2)) & 0x249249249249ULL;
    return x;
}

__global__ void lbvh_keys_kernel(LbvhArgs L) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= L.n_keys) return;
    if (t >= L.m) { L.keys[t] = ~0ULL; return; }
    const int id = L.ids[t];
    const double4 g = L.geo[id];
    const double c[3] = {g.x, g.y, g.z};
    unsigned q[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        double u = (c[k] - L.lo[k]) * L.inv_extent[k] * 65536.0;
        q[k] = u >= 65535.0 ? 65535u : (u > 0.0 ? (unsigned)u : 0u);
    }
    const unsigned long long morton = (lbvh_spread16(q[0]) << 2) | (lbvh_spread16(q[1]) << 1) | lbvh_spread16(q[2]);
    L.keys[t] = (morton << 16) | (unsigned long long)(unsigned)id;
}

__global__ void lbvh_bitonic_kernel(unsigned long long* keys, int n, int j, int k) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int ixj = i ^ j;
    if (ixj > i) {
        unsigned long long a = keys[i], b = keys[ixj];
        bool asc = (i & k) == 0;
        if ((a > b) == asc) { keys[i] = b; keys[ixj] = a; }
    }
}

__device__ __forceinline__ int lbvh_delta(const unsigned long long* keys, int m, int a, int b) {
    if (b < 0 || b >= m) return -1;
    return __clzll((long long)(keys[a] ^ keys[b]));
}

__global__ void lbvh_tree_kernel(LbvhArgs L) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int m = L.m;
    if (i >= m - 1) return;
    const unsigned long long* keys = L.keys;
    const int d = lbvh_delta(keys, m, i, i + 1) - lbvh_delta(keys, m, i, i - 1) >= 0 ? 1 : -1;
    const int dmin = lbvh_delta(keys, m, i, i - d);
    int lmax = 2;
    while (lbvh_delta(keys, m, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (lbvh_delta(keys, m, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = lbvh_delta(keys, m, i, j);
    int s = 0, t = l;
    do {
        t = (t + 1) / 2;
        if (lbvh_delta(keys, m, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const int gamma = i + s * d + (d < 0 ? d : 0);
    const int first = i < j ? i : j, last = i < j ? j : i;
    const int left = first == gamma ? (m - 1) + gamma : gamma;
    const int right = last == gamma + 1 ? (m - 1) + gamma + 1 : gamma + 1;
    L.range[i] = make_int2(first, last);
    L.child[i] = make_int2(left, right);
    L.parent[left] = i;
    L.parent[right] = i;
    if (i == 0) L.parent[0] = -1;
}

__global__ void lbvh_boxes_kernel(LbvhArgs L) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int m = L.m;
    if (k >= m) return;
    const int id = (int)(L.keys[k] & 0xFFFFULL);
    L.leaf_ids[k] = id;
    BvhNode nd;
    {
        const double4 g = L.geo[id];
        const double c[3] = {g.x, g.y, g.z};
        const double r = fabs(L.radius[id]);
#pragma unroll
        for (int a = 0; a < 3; a++) {  // pad outwards: 2^-30 relative (+ tiny absolute), as the host build does
            double lo = c[a] - r, hi = c[a] + r;
            double mm = fmax(fabs(lo), fabs(hi)) + 1.0;
            nd.lo[a] = lo - mm * 9.4e-10; nd.hi[a] = hi + mm * 9.4e-10;
        }
        nd.left = -(k + 1); nd.right = 1; nd.axis = 0; nd.pad = 0;
    }
    int node = (m - 1) + k;
    L.nodes[node] = nd;
    L.depth[node] = 1;
    if (m == 1) { L.nodes[0] = nd; L.depth[0] = 1; return; }
    for (;;) {
        const int p = L.parent[node];
        if (p < 0) break;
        __threadfence();
        if (atomicAdd(&L.visits[p], 1) == 0) break;  // the sibling subtree is not finished: its last thread carries on
        __threadfence();
        const int2 ch = L.child[p];
        const volatile BvhNode* a = L.nodes + ch.x;
        const volatile BvhNode* b = L.nodes + ch.y;
        BvhNode u;
        double ca[3], cb[3];
#pragma unroll
        for (int x = 0; x < 3; x++) {
            const double alo = a->lo[x], ahi = a->hi[x], blo = b->lo[x], bhi = b->hi[x];
            u.lo[x] = fmin(alo, blo); u.hi[x] = fmax(ahi, bhi);
            ca[x] = alo + ahi; cb[x] = blo + bhi;
        }
        int axis = 0;
        if (fabs(ca[1] - cb[1]) > fabs(ca[axis] - cb[axis])) axis = 1;
        if (fabs(ca[2] - cb[2]) > fabs(ca[axis] - cb[axis])) axis = 2;
        const bool swap = ca[axis] > cb[axis];  // "left" is the lower side along the axis (traversal visits the nearer child first)
        const int2 rg = L.range[p];
        const volatile int* dep = L.depth;
        const int da = dep[ch.x], db = dep[ch.y];
        if (rg.y - rg.x + 1 <= 4) { u.left = -(rg.x + 1); u.right = rg.y - rg.x + 1; u.axis = 0; L.depth[p] = 1; }
        else { u.left = swap ? ch.y : ch.x; u.right = swap ? ch.x : ch.y; u.axis = axis; L.depth[p] = 1 + (da > db ? da : db); }
        u.pad = 0;
        L.nodes[p] = u;
        node = p;
    }
}

}  // namespace tray
