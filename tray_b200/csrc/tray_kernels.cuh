// tray_kernels.cuh -- the sm_100a kernels of the tray hot path.
//
//  trace_kernel      persistent megakernel: one path per lane, brute-force closest hit over the
//                    SoA sphere table staged in shared memory, deferred candidate resolution,
//                    lane-level path regeneration from a warp-private pool fed by one global
//                    atomic counter. Replaces the per-sample body of Tracer.RenderLines
//                    (ray/tracer.go:133-143) and everything below it.
//  resolve_kernel    in-order sample sum, *(1/N), sRGB, RGBA8 store (ray/tracer.go:143-152).
//  reference_stream_kernel  conformance mode: one WARP per reference chunk, the chunk's single
//                    sequential stream (ray/tracer.go:120-155), warp-cooperative intersection.
//  first_hit_kernel  RNG-free parity probe of Scene.Hit.
#pragma once
#include "tray_device.cuh"

namespace tray {

constexpr int kMaxDepth = 256;   // attenuation-stack capacity (ids); tray_render rejects larger MaxDepth
#ifndef TRAY_CH
#define TRAY_CH 8
#endif
constexpr int kCand = TRAY_CH + 8;  // deferred candidates per lane before an in-loop flush
constexpr int kBatch = 128;      // samples a warp takes from the global counter at a time
constexpr int kBatchSmall = 32;  // ... in the instantiation for passes with few samples per warp (config 1: 380): the warps run dry together
constexpr unsigned kFull = 0xffffffffu;

// Bounds-check build (-DTRAY_BOUNDS_CHECK, tools/gpu_bounds.sh): every table / list / stack / scratch access of the trace
// kernels is checked against its allocation and violations are counted in stats[6] (tray_stats.bounds_violations) instead of
// being executed. compute-sanitizer is not available on the GPU pool this was developed on, so this build -- run over
// tools/sanitize_run.py, which covers every kernel variant and every table-size remainder -- is the memory-safety evidence;
// the release build compiles the checks away.
#ifdef TRAY_BOUNDS_CHECK
__device__ unsigned long long g_bounds_violations;
#define TRAY_CHECK(cond) ((cond) ? true : (atomicAdd(&g_bounds_violations, 1ull), false))
__device__ __forceinline__ unsigned dyn_smem_end() {  // one past the last byte of the CTA's dynamic shared memory
    unsigned size, base;
    asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(size));
    extern __shared__ __align__(16) unsigned char smem_probe[];
    base = (unsigned)__cvta_generic_to_shared(smem_probe);
    return base + size;
}
#else
#define TRAY_CHECK(cond) true
#endif

struct TraceArgs {
    DevCamera cam;
    int width, spp, max_depth;
    double ray_radius;
    unsigned long long seed;
    // work mapping: local pixel lp -> (x, local row) -> global row
    unsigned long long n_samples;    // samples in this pass
    unsigned long long pass_pixel0;  // first local pixel of this pass
    int row0;                        // params.y0
    int band_rows, shard_count, shard_index;
    int spp_local, sample_stride, sample_offset;  // sample-split: s = sample_offset + j*sample_stride
    unsigned long long* counter;
    double* scratch;                 // n_samples x 3 linear colours, [pixel][sample][rgb]
    unsigned long long* stats;       // [0] segments  [1] depth exhausted
    unsigned long long* progress;    // samples finished (for tray_progress), may be null
    unsigned short* stk_g;           // regroup layout: attenuation stacks, [level][slot] (slot = launch-wide lane id)
    unsigned n_slots;
    const struct CamRay* gen;        // camera rays of the pass, made by camera_ray_kernel
};

#ifndef TRAY_FILTER_PIPE
#define TRAY_FILTER_PIPE 1  // 1: half-chunk double buffer (8 float4 in flight, 128 registers); 0: pair-by-pair (for 96-register builds)
#endif
#ifndef TRAY_PARK_STATE
#define TRAY_PARK_STATE 1  // park the fp64 ray, generator and path counters in shared memory across the filter scan: no spills at 128 registers, 94.7 vs 95.8 ms
#endif
#ifndef TRAY_GEN_INKERNEL
#define TRAY_GEN_INKERNEL 1
#endif
#ifndef TRAY_FP32_SKIP_ORIGIN
// fp32 fast path: a ray that leaves a sphere to its outside cannot meet that sphere again (convex), but in float32 0.8 % of such
// tests report a hit beyond FrontEpsilon (the r = 1000 ground: c = |C-O|^2 - r^2 cancels to +-0.06; short scattered directions). It
// is a back-face hit: the normal flips inwards and the path bounces INSIDE the ground until the depth limit -- 8 % more ray
// segments in all, dark speckles. 1 = the exact test skips the sphere of origin for outward rays (in strict fp64 that test never
// hits, tests/test_fp32_skip_origin_model.py; the fp64 modes follow the reference and test everything).
#define TRAY_FP32_SKIP_ORIGIN 1
#endif
#ifndef TRAY_FP32_FORWARD
// fp32 fast path: the attenuation product is carried forward along the path ((a1*a2)*a3 ... * sky) instead of the reference's
// product on unwind (a1*(a2*(a3*sky)), objects.go:56): the same product up to float32 rounding, no id stack and no unwind loop
// (two dependent loads per level for the 4-5 lanes that finish in an iteration: 6 % of the fp32 kernel's stall samples).
#define TRAY_FP32_FORWARD 1
#endif
// TRAY_GEN_INKERNEL (default): every warp generates the camera rays of its next 32 samples itself -- all 32 lanes busy --
// into a shared-memory pool, and regeneration takes them from there (measured 95.9 ms per config-2 frame). With 0 a separate
// camera_ray_kernel makes all camera rays of the pass ahead and they travel through HBM, 64 B/path each way (97.8 ms).
// Either way the per-sample camera code never runs on the handful of lanes whose path has just ended (101.2 ms).
constexpr unsigned kGenSlots = 32;  // camera rays a warp generates at a time (its pool: 32 slots of 64 bytes)
template <int TPB> constexpr size_t kGenPoolBytes = TRAY_GEN_INKERNEL ? (size_t)(TPB / 32) * kGenSlots * 64 : 0;

// One generated camera ray: what Tracer.RenderLines + Camera.GetRay leave behind for a sample (ray/tracer.go:133-141,
// ray/camera.go:113-142): origin, direction and the generator state after the pixel-jitter and aperture draws.
struct __align__(16) CamRay {
    double o[3], d[3];
    unsigned long long rng_hi, rng_lo;
};

// local sample of the pass -> pixel, sample number, stream; InDisc(RayRadius) iff NumRaysPerPixel > 1; GetRay
__device__ __forceinline__ void camera_sample(const TraceArgs& A, unsigned my_li, Pcg& rng, V3<double>& o64, V3<double>& d64) {
    unsigned lp = (unsigned)A.pass_pixel0 + my_li / (unsigned)A.spp_local;
    int j = (int)(my_li % (unsigned)A.spp_local);
    int s = A.sample_offset + j * A.sample_stride;
    int ly = (int)(lp / (unsigned)A.width);
    int x = (int)(lp - (unsigned)ly * (unsigned)A.width);
    int y = A.shard_count <= 1 ? A.row0 + ly
                               : A.row0 + ((ly / A.band_rows) * A.shard_count + A.shard_index) * A.band_rows + (ly - (ly / A.band_rows) * A.band_rows);
    unsigned long long idx = ((unsigned long long)y * (unsigned)A.width + (unsigned)x) * (unsigned)A.spp + (unsigned)s;
    rng = pcg_new_idx(idx, A.seed);
    double jx = 0.0, jy = 0.0;
    if (A.spp > 1) pcg_in_disc(rng, A.ray_radius, jx, jy, A.cam.indisc_variant);  // ray/tracer.go:136-139
    get_ray(A.cam, rng, (double)x, (double)y, jx, jy, o64, d64);
}

// All camera rays of a pass, one sample per thread with every lane busy: inside the megakernel the same code ran for the
// few lanes of a warp whose path had just ended. 64 bytes per sample written here and read back once by trace_kernel.
__global__ void __launch_bounds__(256) camera_ray_kernel(const __grid_constant__ TraceArgs A, CamRay* __restrict__ out) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n_samples) return;
    Pcg rng;
    V3<double> o64, d64;
    camera_sample(A, (unsigned)i, rng, o64, d64);
    double2* q = reinterpret_cast<double2*>(out + i);
    q[0] = make_double2(o64.x, o64.y); q[1] = make_double2(o64.z, d64.x); q[2] = make_double2(d64.y, d64.z);
    reinterpret_cast<ulonglong2*>(q)[3] = make_ulonglong2(rng.hi, rng.lo);
}

__device__ __forceinline__ int map_row(const TraceArgs& A, int ly) {
    if (A.shard_count <= 1) return A.row0 + ly;
    int band = ly / A.band_rows, within = ly - band * A.band_rows;
    return A.row0 + (band * A.shard_count + A.shard_index) * A.band_rows + within;
}

template <typename T>
__device__ __forceinline__ T t_inf();
template <> __device__ __forceinline__ double t_inf<double>() { return __longlong_as_double(0x7ff0000000000000LL); }
template <> __device__ __forceinline__ float t_inf<float>() { return __int_as_float(0x7f800000); }

// Where the hot loop reads the sphere table from.
//   kGeoParam  : the kernel-parameter constant bank (<= kParamSpheres spheres). The loop index is warp-uniform,
//                so the loads go through the uniform datapath (ULDC) into uniform registers and the FP64
//                instructions take them as operands: no LDS, no vector registers for sphere data.
//   kGeoShared : SoA table staged in shared memory, broadcast LDS.128.
//   kGeoGlobal : read-only global loads (scenes too large for shared memory).
//   kGeoFilter : linear scan behind the exact fp32 pair pre-filter (filter_scan).
//   kGeoBVH    : per-lane BVH traversal (large scenes).
//   kGeoCluster: two-level boxes over chunks of 8 spheres, warp-uniform, then the pair pre-filter on the marked chunks (cluster_scan).
//   kGeoClusterBig: the same walk with a third level of boxes on top and the tables in global memory (cluster_scan_big).
enum { kGeoGlobal = 0, kGeoShared = 1, kGeoParam = 2, kGeoFilter = 3, kGeoBVH = 4, kGeoCluster = 5, kGeoClusterBig = 6 };
#ifndef TRAY_PARAM_GEO
#define TRAY_PARAM_GEO 0
#endif
constexpr int kParamSpheres = 960;  // 960 * 32 B = 30 KB of the 32 KB parameter space

template <typename T, int GEO> struct GeoArg { char unused; };
template <typename T> struct GeoArg<T, kGeoParam> { typename Vec4T<T>::type g[kParamSpheres]; };

// One lane's deferred candidates: sphere ids in index order, in shared memory (cand[k*TPB]).
// Resolve them with the reference's running-closest rule (Scene.Hit, ray/objects.go:37-46): a later sphere
// wins only if strictly closer. Sphere data comes from global memory here (rare path, L1/L2 resident).
template <typename T, bool FMA, int TPB>
__device__ __forceinline__ void resolve_candidates(const typename Vec4T<T>::type* __restrict__ ggeo, const uint16_t* cand, int ncand,
                                                   T ox, T oy, T oz, T dx, T dy, T dz, T a, T& best_t, int& best) {
    const T tmin = front_epsilon<T>();
    for (int k = 0; k < ncand; k++) {
        int id = cand[k * TPB];
        typename Vec4T<T>::type g = ggeo[id];
        T h, c, disc, root;
        sphere_terms<T, FMA>(ox, oy, oz, dx, dy, dz, a, g.x, g.y, g.z, g.w, h, c, disc);
        if (certainly_missed(h, c, disc)) continue;  // false positives of the fp32 pre-filter: exact, no sqrt/div
        // Sphere.Hit, objects.go:90-97. A division whose quotient certainly falls outside (tmin, best_t) is skipped:
        // fl(x/a) >= best_t whenever x >= best_t*a*(1+2^-50), fl(x/a) <= tmin whenever x <= tmin*a*(1-2^-50)
        // (a > 0; one rounding in the product, one in the quotient). The accepted root is computed exactly as written.
        T sq = tsqrt(disc);
        const T up = sizeof(T) == 8 ? T(1.0000000000000009) : T(1.000001), dn = sizeof(T) == 8 ? T(0.9999999999999991) : T(0.999999);
        const T hi = best_t * a * up, lo = tmin * a * dn;
        T x = h - sq;
        bool ok = false;
        if (x < hi && x > lo) { root = tdiv(x, a); ok = root > tmin && root < best_t; }
        if (!ok) {
            x = h + sq;
            if (x < hi && x > lo) { root = tdiv(x, a); ok = root > tmin && root < best_t; }
        }
        if (ok) { best_t = root; best = id; }
    }
}

// Rare path: turn a chunk's candidate bit mask into list entries (bit 7-u <-> sphere base+u), flushing the
// list through the full test when it is about to overflow.
template <typename T, bool FMA, int TPB, int CH>
__device__ __noinline__ void push_candidates(const typename Vec4T<T>::type* __restrict__ ggeo, uint16_t* cand, int* ncand_io,
                                             unsigned mask, int base, T ox, T oy, T oz, T dx, T dy, T dz, T a,
                                             T* best_t, int* best) {
    int ncand = *ncand_io;
    if (ncand > kCand - CH) {
        T bt = *best_t; int b = *best;
        resolve_candidates<T, FMA, TPB>(ggeo, cand, ncand, ox, oy, oz, dx, dy, dz, a, bt, b);
        *best_t = bt; *best = b;
        ncand = 0;
    }
    while (mask) {
        int bit = 31 - __clz(mask);
        if (TRAY_CHECK(ncand < kCand)) cand[ncand * TPB] = (uint16_t)(base + (CH - 1 - bit));
        ncand++;
        mask &= ~(1u << bit);
    }
    *ncand_io = ncand;
}

// Exact Sphere.Hit for order-independent traversal: the root the reference would accept for this sphere
// (objects.go:90-97: near root if it is > tmin, else the far root if that is > tmin) competes with the best so far
// by (t, id) lexicographic order, which is what the reference's "strictly closer wins" scan in index order yields.
template <typename T, bool FMA>
__device__ __forceinline__ void exact_test_unordered(const typename Vec4T<T>::type* __restrict__ ggeo, int id, T ox, T oy, T oz,
                                                     T dx, T dy, T dz, T a, T& best_t, int& best) {
    typename Vec4T<T>::type g = ggeo[id];
    T h, c, disc;
    sphere_terms<T, FMA>(ox, oy, oz, dx, dy, dz, a, g.x, g.y, g.z, g.w, h, c, disc);
    if (certainly_missed(h, c, disc)) return;
    if (disc < T(0)) return;
    const T tmin = front_epsilon<T>();
    T sq = tsqrt(disc);
    T root = tdiv(h - sq, a);
    if (!(root > tmin)) {
        root = tdiv(h + sq, a);
        if (!(root > tmin)) return;
    }
    if (root < best_t || (root == best_t && id < best)) { best_t = root; best = id; }
}

// Closest hit through the BVH. Culling is conservative (padded fp64 boxes, slack on the slab comparison), so every
// sphere the exact test could accept is reached; the result equals the linear scan bit for bit.
template <typename T, bool FMA>
__device__ __forceinline__ void bvh_closest_hit(const DevScene<T>& S, const typename Vec4T<T>::type* __restrict__ ggeo, bool has,
                                                T ox, T oy, T oz, T dx, T dy, T dz, T a, T& best_t, int& best, unsigned& ntests) {
    if (!has) return;
    ntests += S.bvh_n_always;
    for (int k = 0; k < S.bvh_n_always; k++) exact_test_unordered<T, FMA>(ggeo, S.bvh_always[k], ox, oy, oz, dx, dy, dz, a, best_t, best);
    if (!S.bvh) return;
    const double o[3] = {(double)ox, (double)oy, (double)oz};
    const double inv[3] = {1.0 / (double)dx, 1.0 / (double)dy, 1.0 / (double)dz};
    // an origin absurdly far from the scene would make (lo - o) lose all its bits: then do not cull at all
    const bool cull = fmax(fabs(o[0]), fmax(fabs(o[1]), fabs(o[2]))) < 1048576.0 * (S.bvh_extent + 1.0);
    int stack[48];
    int sp = 0;
    bool overflow = false;
    stack[sp++] = 0;
    while (sp > 0) {
        const BvhNode nd = S.bvh[stack[--sp]];
        if (cull) {
            double tn = 0.0, tf = (double)best_t;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                double t1 = (nd.lo[k] - o[k]) * inv[k], t2 = (nd.hi[k] - o[k]) * inv[k];
                tn = fmax(tn, fmin(t1, t2));  // fmin/fmax drop NaNs (0*inf when the origin lies on a slab plane)
                tf = fmin(tf, fmax(t1, t2));
            }
            if (!(tn <= tf * 1.000000001 + 1e-300)) continue;
        }
        if (nd.left < 0) {
            const int first = -nd.left - 1;
            ntests += nd.right;
            for (int k = 0; k < nd.right; k++)
                exact_test_unordered<T, FMA>(ggeo, S.bvh_leaf_ids[first + k], ox, oy, oz, dx, dy, dz, a, best_t, best);
        } else {
            // nearer child last on the stack (popped first) so best_t shrinks early
            const bool neg = (nd.axis == 0 ? (double)dx : (nd.axis == 1 ? (double)dy : (double)dz)) < 0.0;
            if (sp <= 46) {
                stack[sp++] = neg ? nd.left : nd.right;
                stack[sp++] = neg ? nd.right : nd.left;
            } else overflow = true;
        }
    }
    // A tree deeper than the stack (the builders reject those today): never drop a subtree silently -- test every sphere
    // instead (the (t, id) rule makes repeated tests harmless). Slow, correct.
    if (overflow) {
        ntests += S.n;
        for (int i = 0; i < S.n; i++) exact_test_unordered<T, FMA>(ggeo, i, ox, oy, oz, dx, dy, dz, a, best_t, best);
    }
}

// Closest hit of one ray per lane behind the exact fp32 pre-filter: scans the pair table in shared memory, queues the
// survivors in this lane's candidate list (flushing it through the exact test when it is about to overflow). The caller
// resolves what is left in the list. Shared by the megakernel and the wavefront intersect kernel.
template <typename T, bool FMA, int TPB>
__device__ __forceinline__ void filter_scan(const DevScene<T>& S, const float4* sfp, const typename Vec4T<T>::type* __restrict__ ggeo,
                                    uint16_t* cand, bool has, T ox, T oy, T oz, T dx, T dy, T dz, T a,
                                    T& best_t, int& best, int& ncand, unsigned& mask_prev, const volatile T* parked = nullptr) {
    constexpr int CH = TRAY_CH;
    const int n_pad = S.n_pad;
    // ---- fp32 conservative pre-filter (packed FFMA2, two spheres per instruction, 8 instructions per pair) ----
    // A test is skipped only when the fp32 evaluation PROVES that Sphere.Hit returns false in strict fp64
    // ("exact pre-filter"). Per sphere, with d = D/|D|, h~ = d.(C-O), c = |C-O|^2 - r^2 and the
    // table holding (C, -(|C|^2 - r^2)):
    //   hp ~ h~ + EH = d.C - d.O + EH        (3 FMA)      |hp - (h~ + EH)| <= Eh = 16u R < EH,   R = max|C|inf + |O|inf
    //   nc ~ -c + E                          (1 ADD + 3 FMA, expanded as -(|C|^2 - r^2) - |O|^2 + 2 O.C + E)
    //   v1 = hp*hp + nc                      (1 FMA)      errors of nc and of this rounding <= u (22 R^2 + 6.2 max r^2) < E
    //   v1 < 0 :  h~ >= 0  =>  hp >= h~ >= 0, so h~^2 - c <= hp^2 - c < 0: disc < 0 in exact and in fp64 arithmetic
    //             h~ <  0  =>  c > hp^2 + (E - error) > 0: h < 0 < c, both roots <= 0 (or disc < 0)
    //   hp < 0 and nc < 0  =>  h~ < 0 < c as well
    // Everything else (incl. NaN table entries = spheres outside the filter's range, canonical NaN has sign 0)
    // goes to the exact path.
    const float u32 = 5.9604645e-8f;
    const double inv_n = 1.0 / sqrt((double)a);
    const double ddx = (double)dx * inv_n, ddy = (double)dy * inv_n, ddz = (double)dz * inv_n;
    const float fdx = (float)ddx, fdy = (float)ddy, fdz = (float)ddz;
    float ndo = -(float)(ddx * (double)ox + ddy * (double)oy + ddz * (double)oz);
    const float mo = 1.0000002f * fmaxf(fabsf((float)(double)ox), fmaxf(fabsf((float)(double)oy), fabsf((float)(double)oz)));
    const float R = S.filt_mc + mo;
    float eh = 17.5f * u32 * R;                                             // > 1.01 * Eh
    float noot = (float)((double)(1.03f * u32 * (22.0f * R * R + 6.2f * S.filt_r2max) + 1e-30f) -
                         ((double)ox * (double)ox + (double)oy * (double)oy + (double)oz * (double)oz));  // E - |O|^2
    if (!(mo < 1e6f)) { noot = __int_as_float(0x7f800000); eh = 0.0f; }     // origin too far out: filter off for this ray
    ndo += eh;                                                              // one more rounding <= u(|d.O| + EH), inside Eh
    const float2 Dx = make_float2(fdx, fdx), Dy = make_float2(fdy, fdy), Dz = make_float2(fdz, fdz);
    const float px = (float)(2.0 * (double)ox), py = (float)(2.0 * (double)oy), pz = (float)(2.0 * (double)oz);
    const float2 Px = make_float2(px, px), Py = make_float2(py, py), Pz = make_float2(pz, pz);
    const float2 NDO = make_float2(ndo, ndo), NOOT = make_float2(noot, noot);
    // Software-pipelined by half chunks: while the two pairs in registers A are evaluated, the LDS.128 of the next
    // two pairs (B) are in flight, and vice versa; the table is followed by other shared data, so the last prefetch
    // reads valid (unused) memory. One running shared-memory address, immediate offsets.
    static_assert(CH == 8, "the pre-filter loop is written for chunks of 8 spheres");
    auto lds4 = [](float4& v, unsigned addr) {
#ifdef TRAY_BOUNDS_CHECK
        if (!TRAY_CHECK(addr + 16u <= dyn_smem_end())) { v = make_float4(0, 0, 0, 0); return; }  // (the last prefetch reads past the table: it must stay inside the allocation)
#endif
        asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    };
    unsigned mask = 0;
    auto pair = [&](const float4& g0, const float4& g1) {
        const float2 cx = make_float2(g0.x, g0.y), cy = make_float2(g0.z, g0.w), cz = make_float2(g1.x, g1.y);
        float2 h = __ffma2_rn(Dx, cx, __ffma2_rn(Dy, cy, __ffma2_rn(Dz, cz, NDO)));
        float2 nko = __fadd2_rn(make_float2(g1.z, g1.w), NOOT);                    // -(|C|^2 - r^2) - |O|^2 + E
        float2 nc = __ffma2_rn(Px, cx, __ffma2_rn(Py, cy, __ffma2_rn(Pz, cz, nko)));
        float2 v1 = __ffma2_rn(h, h, nc);
        int mx = __float_as_int(v1.x) | (__float_as_int(h.x) & __float_as_int(nc.x));
        int my = __float_as_int(v1.y) | (__float_as_int(h.y) & __float_as_int(nc.y));
        mask = __funnelshift_l((unsigned)mx, mask, 1);
        mask = __funnelshift_l((unsigned)my, mask, 1);
    };
    unsigned saddr = (unsigned)__cvta_generic_to_shared(sfp);
#if TRAY_FILTER_PIPE
    float4 a0, a1, a2, a3, b0, b1, b2, b3;
    lds4(a0, saddr); lds4(a1, saddr + 16); lds4(a2, saddr + 32); lds4(a3, saddr + 48);
#pragma unroll 1
    for (int i = 0; i < n_pad; i += CH) {
        lds4(b0, saddr + 64); lds4(b1, saddr + 80); lds4(b2, saddr + 96); lds4(b3, saddr + 112);
        mask = 0;
        pair(a0, a1); pair(a2, a3);
        lds4(a0, saddr + 128); lds4(a1, saddr + 144); lds4(a2, saddr + 160); lds4(a3, saddr + 176);
        pair(b0, b1); pair(b2, b3);
#else
    // low-register form (more resident warps hide the LDS latency instead): one pair in flight ahead of the one evaluated
    float4 a0, a1, b0, b1;
    lds4(a0, saddr); lds4(a1, saddr + 16);
#pragma unroll 1
    for (int i = 0; i < n_pad; i += CH) {
        mask = 0;
        lds4(b0, saddr + 32); lds4(b1, saddr + 48);
        pair(a0, a1);
        lds4(a0, saddr + 64); lds4(a1, saddr + 80);
        pair(b0, b1);
        lds4(b0, saddr + 96); lds4(b1, saddr + 112);
        pair(a0, a1);
        lds4(a0, saddr + 128); lds4(a1, saddr + 144);
        pair(b0, b1);
#endif
        saddr += 128;
        mask = ~mask & 0xffu;  // 1 = must be tested exactly
        if (mask_prev) {
            if (ncand > kCand - CH) {  // list about to overflow (rare): run the exact test on what is queued
                if (parked) {  // TRAY_PARK_STATE: the fp64 ray waits in shared memory during the scan
                    const T px_ = T(parked[0]), py_ = T(parked[TPB]), pz_ = T(parked[2 * TPB]);
                    const T qx_ = T(parked[3 * TPB]), qy_ = T(parked[4 * TPB]), qz_ = T(parked[5 * TPB]);
                    push_candidates<T, FMA, TPB, CH>(ggeo, cand, &ncand, 0u, 0, px_, py_, pz_, qx_, qy_, qz_, (qx_ * qx_ + qy_ * qy_) + qz_ * qz_, &best_t, &best);
                } else {
                    push_candidates<T, FMA, TPB, CH>(ggeo, cand, &ncand, 0u, 0, ox, oy, oz, dx, dy, dz, a, &best_t, &best);
                }
            }
            unsigned m = mask_prev;
            do {  // append in index order: bit CH-1-u <-> sphere (i-CH)+u
                int bit = 31 - __clz(m);
                if (TRAY_CHECK(ncand < kCand)) cand[ncand * TPB] = (uint16_t)(i - 1 - bit);
                ncand++;
                m &= ~(1u << bit);
            } while (m);
        }
        mask_prev = has ? mask : 0u;  // idle lanes queue nothing
    }
}

// Deferred candidates in ANY order (cluster scan): Sphere.Hit's own root selection (objects.go:90-97) per sphere, the
// winner by (t, id) lexicographic order -- what the reference's index-order scan with "strictly closer wins" yields
// (tests/test_unordered_rule_model.py). Divisions whose quotient certainly loses are skipped as in resolve_candidates;
// the upper bound is strict (fl(x/a) > best_t whenever x >= best_t*a*(1+2^-49)), so a tie with a lower id is never skipped.
template <typename T, bool FMA, int TPB>
__device__ __forceinline__ void resolve_candidates_lex(const typename Vec4T<T>::type* __restrict__ ggeo, const uint16_t* cand, int ncand,
                                                       T ox, T oy, T oz, T dx, T dy, T dz, T a, T& best_t, int& best, int skip = -1) {
    const T tmin = front_epsilon<T>();
    const T ra = trcp(a);  // every root is divided by the same a: one reciprocal chain per ray segment (rcp_refined)
    const T up = sizeof(T) == 8 ? T(1.0000000000000018) : T(1.000002), dn = sizeof(T) == 8 ? T(0.9999999999999991) : T(0.999999);
    const T lo = tmin * a * dn;
    T hi = best_t * a * up;    // follows best_t (updated where a candidate wins)
#pragma unroll 1
    for (int k = 0; k < ncand; k++) {  // (not unrolled: one copy of the exact test in the instruction cache)
        const int id = cand[k * TPB];
        if (!TRAY_CHECK(k < kCand && id >= 0 && id <= 65535)) continue;  // (the table holds n_pad + 8 entries; ids come from the staged tables)
        if constexpr (sizeof(T) == 4 && TRAY_FP32_SKIP_ORIGIN) { if (id == skip) continue; }  // fp32 fast path: the sphere the ray leaves outwards
        typename Vec4T<T>::type g = ggeo[id];
        T h, c, disc, root;
        sphere_terms<T, FMA>(ox, oy, oz, dx, dy, dz, a, g.x, g.y, g.z, g.w, h, c, disc);
        if (certainly_missed(h, c, disc)) continue;
        T sq = tsqrt_hot(disc);
        T x = h - sq;
        bool ok = false;
        if (x < hi && x > lo) { root = tdiv_r(x, a, ra); ok = root > tmin && (root < best_t || (root == best_t && id < best)); }
        if (!ok) {
            x = h + sq;
            if (x < hi && x > lo) { root = tdiv_r(x, a, ra); ok = root > tmin && (root < best_t || (root == best_t && id < best)); }
        }
        if (ok) { best_t = root; best = id; hi = root * a * up; }
    }
}

// The rare in-scan flush of a full candidate list: one out-of-line copy (results in registers), so that the scan's hot
// code stays small.
template <typename T> struct BestHit { T t; int id; };
template <typename T, bool FMA, int TPB>
__device__ __noinline__ BestHit<T> flush_candidates_lex(const typename Vec4T<T>::type* __restrict__ ggeo, const uint16_t* cand, int ncand,
                                                        T ox, T oy, T oz, T dx, T dy, T dz, T best_t, int best) {
    resolve_candidates_lex<T, FMA, TPB>(ggeo, cand, ncand, ox, oy, oz, dx, dy, dz, (dx * dx + dy * dy) + dz * dz, best_t, best);
    BestHit<T> r;
    r.t = best_t; r.id = best;
    return r;
}

// Two-level closest hit (kGeoCluster, the default for scenes that fit): the spheres are grouped at upload into spatially
// compact chunks of 8 (recursive median split), 8 chunks form a group. Per ray segment:
//   1. the boxes of the groups are tested, 8 at a time, with a conservative fp32 slab test (packed FFMA2/FADD2 + FMNMX3);
//   2. the warp takes the UNION of its lanes' "may hit" bits (REDUX.OR), so everything below stays warp-uniform: table
//      loads are broadcast LDS.128, no lane waits for another;
//   3. per marked group the boxes of its 8 chunks are tested the same way, union again;
//   4. only the marked chunks run the exact fp32 pair pre-filter of filter_scan; survivors are queued and resolved in
//      fp64 by (t, id) lexicographic order (= the reference's scan order result).
// Spheres the filter cannot bound (|C| or r^2 > 256, non-finite) and very large ones sit in "always" groups that every
// ray scans. On the benchmark scene a warp scans about 5 of 62 chunks per segment (tools/cluster_sim.py).
//
// The slab test of a box (centre c, half extent e, fp32) for the ray O + d t, d = Unit(D), per axis k:
//   inv = rcp(fl32(d_k)) (MUFU.RCP, |d_k| clamped to >= 2^-60), nq = -fl32(O_k*inv), ai = |inv|, sl = 16 u R ai   (per ray;
//                                                                     u = 2^-24, R = max|box coordinate| + |O|inf)
//   A = fma(c, inv, nq)   B = fma(e, ai, sl)   near = A - B   far = A + B
//   miss  <=>  min_k far < max_k near  or  min_k far < 0          (sign bits of two words, one LOP3)
// If the exact ray meets the box at some t* >= 0 then for every axis |(c-O)inv - t*| <= e ai + 3.03 u R ai (inv is the exact
// reciprocal of a direction within 3u of d_k: one rounding of d_k to fp32, 2u of the approximate reciprocal), A is off by
// <= 2.01 u R ai (rounding of O*inv and of the fma), B may fall short by u R ai, near/far round by <= 2 u R ai: 8.1 u R ai
// in all. A hit the strict fp64 Sphere.Hit reports lies within sqrt(160 eps) R = 1.33e-7 R = 2.2 u R of the sphere along the
// true ray (its discriminant is exact to 20 eps a (|C-O|^2 + r^2)), and the boxes are padded on top of that: 10.3 < 16, so
// near_k <= t* <= far_k for every k and the test cannot report a miss. tests/test_cluster_box_model.py replays the test in
// exact single-rounded arithmetic with the reciprocal pushed to either end of its error interval.
template <typename T, bool FMA, int TPB>
__device__ __forceinline__ void cluster_scan(const DevScene<T>& S, const float4* sblob, const typename Vec4T<T>::type* __restrict__ ggeo,
                                             uint16_t* cand, bool has, T ox, T oy, T oz, T dx, T dy, T dz, T a, T udx, T udy, T udz,
                                             T& best_t, int& best, int& ncand, unsigned& nchunks, unsigned& nboxes,
                                             const volatile T* parked = nullptr) {
    // ---- per-ray constants of the pair pre-filter (as filter_scan; the unit direction is Unit(D) of the RayColor step,
    //      computed by the caller in fp64 with IEEE sqrt and divisions: within 2 ulp of D/|D|, far inside the bounds) ----
    const float u32 = 5.9604645e-8f;
    const double ddx = (double)udx, ddy = (double)udy, ddz = (double)udz;
    const float fdx = (float)ddx, fdy = (float)ddy, fdz = (float)ddz;
    float ndo = -(float)(ddx * (double)ox + ddy * (double)oy + ddz * (double)oz);
    const float mo = 1.0000002f * fmaxf(fabsf((float)(double)ox), fmaxf(fabsf((float)(double)oy), fabsf((float)(double)oz)));
    const float R = S.cl_r + mo;
    float eh = 17.5f * u32 * R;
    float noot = (float)((double)(1.03f * u32 * (22.0f * R * R + 6.2f * S.filt_r2max) + 1e-30f) -
                         ((double)ox * (double)ox + (double)oy * (double)oy + (double)oz * (double)oz));
    // origin too far out, or a direction without a finite positive length: no culling at all for this ray
    const bool off = !(mo < 1e6f) || !((double)a > 0.0 && (double)a < 1.7976931348623157e308);
    if (off) { noot = __int_as_float(0x7f800000); eh = 0.0f; }
    ndo += eh;
    const float2 Dx = make_float2(fdx, fdx), Dy = make_float2(fdy, fdy), Dz = make_float2(fdz, fdz);
    const float px = (float)(2.0 * (double)ox), py = (float)(2.0 * (double)oy), pz = (float)(2.0 * (double)oz);
    const float2 Px = make_float2(px, px), Py = make_float2(py, py), Pz = make_float2(pz, pz);
    const float2 NDO = make_float2(ndo, ndo), NOOT = make_float2(noot, noot);
    // ---- per-ray constants of the box test ----
    auto rcp = [](float d) {  // MUFU.RCP: within 1 ulp (2u) of 1/d; the operand is a normal number and so is the result
        const float lim = 8.6736174e-19f;  // 2^-60
        float r;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fabsf(d) < lim ? copysignf(lim, d) : d));
        return r;
    };
    const float ix = rcp(fdx), iy = rcp(fdy), iz = rcp(fdz);
    const float nqx = -(float)((double)ox * (double)ix), nqy = -(float)((double)oy * (double)iy), nqz = -(float)((double)oz * (double)iz);
    const float ks = 16.0f * u32 * R;
    const float2 IX = make_float2(ix, ix), IY = make_float2(iy, iy), IZ = make_float2(iz, iz);
    const float2 NQX = make_float2(nqx, nqx), NQY = make_float2(nqy, nqy), NQZ = make_float2(nqz, nqz);
    const float2 AX = make_float2(fabsf(ix), fabsf(ix)), AY = make_float2(fabsf(iy), fabsf(iy)), AZ = make_float2(fabsf(iz), fabsf(iz));
    const float2 SX = make_float2(ks * fabsf(ix), ks * fabsf(ix)), SY = make_float2(ks * fabsf(iy), ks * fabsf(iy)),
                 SZ = make_float2(ks * fabsf(iz), ks * fabsf(iz));

#ifdef TRAY_BOUNDS_CHECK
    const unsigned chk_lo = (unsigned)__cvta_generic_to_shared(sblob), chk_hi = chk_lo + (unsigned)S.cl_blob_f4 * 16u;
    auto lds4 = [&](float4& v, unsigned addr) {  // every table read of the scan stays inside the staged blob
        if (!TRAY_CHECK(addr >= chk_lo && addr + 16u <= chk_hi && addr + 16u <= dyn_smem_end())) { v = make_float4(0, 0, 0, 0); return; }
        asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    };
#else
    auto lds4 = [](float4& v, unsigned addr) {
        asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    };
#endif
    // Every pair of tests yields its own 2-bit word and the four words of a chunk are merged by a tree, so that the bit
    // gathering is a dependency chain of 4 instructions instead of 8 funnel shifts in a row (the scan is latency-bound).
    auto pair = [&](const float4& g0, const float4& g1) -> unsigned {  // two spheres: 8 packed instructions (see filter_scan)
        const float2 cx = make_float2(g0.x, g0.y), cy = make_float2(g0.z, g0.w), cz = make_float2(g1.x, g1.y);
        float2 h = __ffma2_rn(Dx, cx, __ffma2_rn(Dy, cy, __ffma2_rn(Dz, cz, NDO)));
        float2 nko = __fadd2_rn(make_float2(g1.z, g1.w), NOOT);
        float2 nc = __ffma2_rn(Px, cx, __ffma2_rn(Py, cy, __ffma2_rn(Pz, cz, nko)));
        float2 v1 = __ffma2_rn(h, h, nc);
        int mx = __float_as_int(v1.x) | (__float_as_int(h.x) & __float_as_int(nc.x));
        int my = __float_as_int(v1.y) | (__float_as_int(h.y) & __float_as_int(nc.y));
        return __funnelshift_l((unsigned)my, (unsigned)mx >> 31, 1);   // bit 1: first sphere missed, bit 0: second
    };
    auto merge4 = [](unsigned p0, unsigned p1, unsigned p2, unsigned p3) -> unsigned { return (((p0 << 2) | p1) << 4) | ((p2 << 2) | p3); };
    auto boxes = [&](unsigned addr) -> unsigned {  // two boxes: 12 packed instructions + 2 x (2 FMNMX3 + FADD + LOP3 + SHF)
        float4 b0, b1, b2;
        lds4(b0, addr); lds4(b1, addr + 16); lds4(b2, addr + 32);
        const float2 ax = __ffma2_rn(make_float2(b0.x, b0.y), IX, NQX), ay = __ffma2_rn(make_float2(b0.z, b0.w), IY, NQY),
                     az = __ffma2_rn(make_float2(b1.x, b1.y), IZ, NQZ);
        const float2 bx = __ffma2_rn(make_float2(b1.z, b1.w), AX, SX), by = __ffma2_rn(make_float2(b2.x, b2.y), AY, SY),
                     bz = __ffma2_rn(make_float2(b2.z, b2.w), AZ, SZ);
        const float2 nx = __fadd2_rn(ax, make_float2(-bx.x, -bx.y)), ny = __fadd2_rn(ay, make_float2(-by.x, -by.y)),
                     nz = __fadd2_rn(az, make_float2(-bz.x, -bz.y));
        const float2 fx = __fadd2_rn(ax, bx), fy = __fadd2_rn(ay, by), fz = __fadd2_rn(az, bz);
        const float tn0 = fmaxf(fmaxf(nx.x, ny.x), nz.x), tf0 = fminf(fminf(fx.x, fy.x), fz.x);
        const float tn1 = fmaxf(fmaxf(nx.y, ny.y), nz.y), tf1 = fminf(fminf(fx.y, fy.y), fz.y);
        return __funnelshift_l((unsigned)(__float_as_int(tf1 - tn1) | __float_as_int(tf1)),
                               (unsigned)(__float_as_int(tf0 - tn0) | __float_as_int(tf0)) >> 31, 1);
    };
    auto may_hit8 = [&](unsigned mask) -> unsigned {  // 8 "missed" bits -> this lane's "may hit" byte, then the warp's union
        unsigned hit = off ? 0xffu : (~mask & 0xffu);
        if (!has) hit = 0u;
        return __reduce_or_sync(kFull, hit);
    };
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(sblob);
    const unsigned a_box2 = sbase + (unsigned)S.cl_off_box2 * 16u, a_box1 = sbase + (unsigned)S.cl_off_box1 * 16u;
    const uint16_t* sids = reinterpret_cast<const uint16_t*>(sblob + S.cl_off_ids);
    const int n_always = S.cl_always_groups, n_words = S.cl_real_groups >> 3;
    // words -n_always..-1: one always-group each (no box tests); words 0..: 8 real groups = 64 chunks behind their boxes.
    // Per word: (1) the 8 group boxes, warp union; (2) the chunk boxes of every marked group, each lane collecting its own
    // 64-bit chunk mask -- no cross-lane step between groups, so their tests overlap --, then ONE union of the mask; (3) the
    // marked chunks through the pair pre-filter, software-pipelined by half chunks (the loads of the next half / next chunk
    // are in flight while a half is evaluated). The scan is latency-bound: the point of this shape is few serial points.
#pragma unroll 1
    for (int w = -n_always; w < n_words; w++) {
        unsigned long long ucm;  // warp-uniform: chunks of this word to scan, bit 63-p <-> chunk cbase+p
        int cbase;
        if (w < 0) {
            cbase = (S.cl_real_groups + (w + n_always)) * 8;
            ucm = (unsigned long long)(w == -1 ? S.cl_always_last : 0xffu) << 56;
        } else {
            cbase = w * 64;
            // ONE copy of the 8-box test serves both levels (instruction-cache footprint: the trace kernel's per-segment code is
            // at the size of the 32 KB instruction cache): first the word's 8 group boxes (gi < 0), then the 8 chunk boxes of
            // every marked group. The branch on gi is warp-uniform.
            unsigned adb = a_box1 + (unsigned)w * (4u * 48u);
            unsigned ug = 0;
            int gi = -1;
            unsigned long long cm = 0;
#pragma unroll 1
            for (;;) {
                const unsigned m8 = merge4(boxes(adb), boxes(adb + 48), boxes(adb + 96), boxes(adb + 144));
                nboxes += 8;
                if (gi < 0) ug = may_hit8(m8);
                else cm |= (unsigned long long)(off ? 0xffu : (~m8 & 0xffu)) << (56 - 8 * gi);
                if (!ug) break;
                const int gbit = 31 - __clz(ug);
                ug &= ~(1u << gbit);
                gi = 7 - gbit;
                adb = a_box2 + (unsigned)(w * 8 + gi) * (4u * 48u);
            }
            if (!has) cm = 0;
            const unsigned lo = __reduce_or_sync(kFull, (unsigned)cm), hi = __reduce_or_sync(kFull, (unsigned)(cm >> 32));
            ucm = ((unsigned long long)hi << 32) | lo;
        }
        if (ucm == 0) continue;
        int pos = __clzll((long long)ucm);
        unsigned ad = sbase + (unsigned)(cbase + pos) * 128u;
        float4 a0, a1, a2, a3, b0, b1, b2, b3;
        lds4(a0, ad); lds4(a1, ad + 16); lds4(a2, ad + 32); lds4(a3, ad + 48);
#pragma unroll 1
        for (;;) {
            lds4(b0, ad + 64); lds4(b1, ad + 80); lds4(b2, ad + 96); lds4(b3, ad + 112);
            const unsigned long long rest = ucm & ~(0x8000000000000000ull >> pos);
            const unsigned p0 = pair(a0, a1), p1 = pair(a2, a3);
            const int npos = rest ? __clzll((long long)rest) : pos;   // (last chunk: the prefetch re-reads it, harmlessly)
            const unsigned nad = sbase + (unsigned)(cbase + npos) * 128u;
            lds4(a0, nad); lds4(a1, nad + 16); lds4(a2, nad + 32); lds4(a3, nad + 48);
            const unsigned mask = merge4(p0, p1, pair(b0, b1), pair(b2, b3));
            nchunks++;
            unsigned m = has ? (~mask & 0xffu) : 0u;  // 1 = must be tested exactly
            if (m) {
                const int chunk = cbase + pos;
                if (ncand > kCand - 8) {  // list about to overflow (rare): run the exact test on what is queued
                    T fx_ = ox, fy_ = oy, fz_ = oz, gx_ = dx, gy_ = dy, gz_ = dz;
                    if (parked) {  // TRAY_PARK_STATE: the fp64 ray waits in shared memory during the scan
                        fx_ = T(parked[0]); fy_ = T(parked[TPB]); fz_ = T(parked[2 * TPB]);
                        gx_ = T(parked[3 * TPB]); gy_ = T(parked[4 * TPB]); gz_ = T(parked[5 * TPB]);
                    }
                    const BestHit<T> bh = flush_candidates_lex<T, FMA, TPB>(ggeo, cand, ncand, fx_, fy_, fz_, gx_, gy_, gz_, best_t, best);
                    best_t = bh.t; best = bh.id;
                    ncand = 0;
                }
                do {  // bit 7-u <-> slot chunk*8+u
                    const int bit = 31 - __clz(m);
#ifdef TRAY_BOUNDS_CHECK
                    const bool ids_inside = (unsigned)__cvta_generic_to_shared(sids + (chunk * 8 + 7 - bit)) + 2u <= chk_hi;
#else
                    const bool ids_inside = true;
                    (void)ids_inside;
#endif
                    if (TRAY_CHECK(ids_inside && ncand < kCand && chunk * 8 + 7 - bit < (S.cl_off_box2 >> 3) * 8 && chunk >= 0))
                        cand[ncand * TPB] = sids[chunk * 8 + 7 - bit];
                    ncand++;
                    m &= ~(1u << bit);
                } while (m);
            }
            if (!rest) break;
            ucm = rest; pos = npos; ad = nad;
        }
    }
}

// The cluster scan for scenes whose tables do not fit shared memory (kGeoClusterBig: 2049 .. 32768 spheres, BASELINE config 4):
// same structure and the same tests as cluster_scan, with
//   * the tables read from GLOBAL memory through the read-only path -- every load of the walk is warp-uniform (one L1
//     transaction, broadcast), the hot part of the 200 KB of tables stays in L1/L2 --, and
//   * a third level on top: one box per WORD of 8 groups (512 slots). The word boxes are tested first, 8 at a time, the
//     warp's union of marked words is one 64-bit mask, and only marked words run the group -> chunk -> pair-filter walk.
//     (cluster_scan tests the group boxes of every word: 8 box tests per 512 spheres for every ray.)
// Conservativeness is that of cluster_scan: a box of any level contains the boxes below it (build_clusters), the slab test
// and its bound are the same code. Results are bit-identical to the linear scan (GPU tests on 10 001 spheres).
template <typename T, bool FMA, int TPB>
__device__ __forceinline__ void cluster_scan_big(const DevScene<T>& S, const typename Vec4T<T>::type* __restrict__ ggeo,
                                                 uint16_t* cand, bool has, T ox, T oy, T oz, T dx, T dy, T dz, T a, T udx, T udy, T udz,
                                                 T& best_t, int& best, int& ncand, unsigned& nchunks, unsigned& nboxes,
                                                 const volatile T* parked = nullptr) {
    // ---- per-ray constants: as cluster_scan ----
    const float u32 = 5.9604645e-8f;
    const double ddx = (double)udx, ddy = (double)udy, ddz = (double)udz;
    const float fdx = (float)ddx, fdy = (float)ddy, fdz = (float)ddz;
    float ndo = -(float)(ddx * (double)ox + ddy * (double)oy + ddz * (double)oz);
    const float mo = 1.0000002f * fmaxf(fabsf((float)(double)ox), fmaxf(fabsf((float)(double)oy), fabsf((float)(double)oz)));
    const float R = S.cl_r + mo;
    float eh = 17.5f * u32 * R;
    float noot = (float)((double)(1.03f * u32 * (22.0f * R * R + 6.2f * S.filt_r2max) + 1e-30f) -
                         ((double)ox * (double)ox + (double)oy * (double)oy + (double)oz * (double)oz));
    const bool off = !(mo < 1e6f) || !((double)a > 0.0 && (double)a < 1.7976931348623157e308);
    if (off) { noot = __int_as_float(0x7f800000); eh = 0.0f; }
    ndo += eh;
    const float2 Dx = make_float2(fdx, fdx), Dy = make_float2(fdy, fdy), Dz = make_float2(fdz, fdz);
    const float px = (float)(2.0 * (double)ox), py = (float)(2.0 * (double)oy), pz = (float)(2.0 * (double)oz);
    const float2 Px = make_float2(px, px), Py = make_float2(py, py), Pz = make_float2(pz, pz);
    const float2 NDO = make_float2(ndo, ndo), NOOT = make_float2(noot, noot);
    auto rcp = [](float d) {
        const float lim = 8.6736174e-19f;  // 2^-60
        float r;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fabsf(d) < lim ? copysignf(lim, d) : d));
        return r;
    };
    const float ix = rcp(fdx), iy = rcp(fdy), iz = rcp(fdz);
    const float nqx = -(float)((double)ox * (double)ix), nqy = -(float)((double)oy * (double)iy), nqz = -(float)((double)oz * (double)iz);
    const float ks = 16.0f * u32 * R;
    const float2 IX = make_float2(ix, ix), IY = make_float2(iy, iy), IZ = make_float2(iz, iz);
    const float2 NQX = make_float2(nqx, nqx), NQY = make_float2(nqy, nqy), NQZ = make_float2(nqz, nqz);
    const float2 AX = make_float2(fabsf(ix), fabsf(ix)), AY = make_float2(fabsf(iy), fabsf(iy)), AZ = make_float2(fabsf(iz), fabsf(iz));
    const float2 SX = make_float2(ks * fabsf(ix), ks * fabsf(ix)), SY = make_float2(ks * fabsf(iy), ks * fabsf(iy)),
                 SZ = make_float2(ks * fabsf(iz), ks * fabsf(iz));

    const float4* __restrict__ gb = S.cl_blob;
    auto ldg4 = [&](float4& v, unsigned idx) {  // idx: float4 index into the blob (warp-uniform)
        if (!TRAY_CHECK(idx < (unsigned)S.cl_blob_f4)) { v = make_float4(0, 0, 0, 0); return; }
        v = __ldg(gb + idx);
    };
    auto pair = [&](const float4& g0, const float4& g1) -> unsigned {  // two spheres: 8 packed instructions (see filter_scan)
        const float2 cx = make_float2(g0.x, g0.y), cy = make_float2(g0.z, g0.w), cz = make_float2(g1.x, g1.y);
        float2 h = __ffma2_rn(Dx, cx, __ffma2_rn(Dy, cy, __ffma2_rn(Dz, cz, NDO)));
        float2 nko = __fadd2_rn(make_float2(g1.z, g1.w), NOOT);
        float2 nc = __ffma2_rn(Px, cx, __ffma2_rn(Py, cy, __ffma2_rn(Pz, cz, nko)));
        float2 v1 = __ffma2_rn(h, h, nc);
        int mx = __float_as_int(v1.x) | (__float_as_int(h.x) & __float_as_int(nc.x));
        int my = __float_as_int(v1.y) | (__float_as_int(h.y) & __float_as_int(nc.y));
        return __funnelshift_l((unsigned)my, (unsigned)mx >> 31, 1);
    };
    auto merge4 = [](unsigned p0, unsigned p1, unsigned p2, unsigned p3) -> unsigned { return (((p0 << 2) | p1) << 4) | ((p2 << 2) | p3); };
    auto boxes = [&](unsigned idx) -> unsigned {  // two boxes (three float4), see cluster_scan
        float4 b0, b1, b2;
        ldg4(b0, idx); ldg4(b1, idx + 1); ldg4(b2, idx + 2);
        const float2 ax = __ffma2_rn(make_float2(b0.x, b0.y), IX, NQX), ay = __ffma2_rn(make_float2(b0.z, b0.w), IY, NQY),
                     az = __ffma2_rn(make_float2(b1.x, b1.y), IZ, NQZ);
        const float2 bx = __ffma2_rn(make_float2(b1.z, b1.w), AX, SX), by = __ffma2_rn(make_float2(b2.x, b2.y), AY, SY),
                     bz = __ffma2_rn(make_float2(b2.z, b2.w), AZ, SZ);
        const float2 nx = __fadd2_rn(ax, make_float2(-bx.x, -bx.y)), ny = __fadd2_rn(ay, make_float2(-by.x, -by.y)),
                     nz = __fadd2_rn(az, make_float2(-bz.x, -bz.y));
        const float2 fx = __fadd2_rn(ax, bx), fy = __fadd2_rn(ay, by), fz = __fadd2_rn(az, bz);
        const float tn0 = fmaxf(fmaxf(nx.x, ny.x), nz.x), tf0 = fminf(fminf(fx.x, fy.x), fz.x);
        const float tn1 = fmaxf(fmaxf(nx.y, ny.y), nz.y), tf1 = fminf(fminf(fx.y, fy.y), fz.y);
        return __funnelshift_l((unsigned)(__float_as_int(tf1 - tn1) | __float_as_int(tf1)),
                               (unsigned)(__float_as_int(tf0 - tn0) | __float_as_int(tf0)) >> 31, 1);
    };
    auto may_hit8 = [&](unsigned mask) -> unsigned {
        unsigned hit = off ? 0xffu : (~mask & 0xffu);
        if (!has) hit = 0u;
        return __reduce_or_sync(kFull, hit);
    };
    const uint16_t* __restrict__ gids = reinterpret_cast<const uint16_t*>(gb + S.cl_off_ids);
    const int n_always = S.cl_always_groups, n_words = S.cl_real_groups >> 3;
    // The walk is a warp-uniform state machine around ONE copy of the 8-box test and ONE copy of the chunk loop:
    //   kL0: 8 word boxes -> uw (marked words, bit 63-w)        kL1: the 8 group boxes of word w -> ug
    //   kL2: the 8 chunk boxes of group gi of word w -> cm      kNextGroup / kNextWord: bookkeeping, no tests
    //   kChunks: the marked chunks (ucm) of the current word through the pair pre-filter
    // Always-words (spheres outside the filter's range, very large ones) are scanned by every ray, after level 0.
    enum { kL0 = 0, kL1 = 1, kL2 = 2, kNextGroup = 3, kNextWord = 4, kChunks = 5 };
    const int n_oct0 = (n_words + 7) >> 3;
    unsigned long long uw = 0, cm = 0, ucm = 0;
    unsigned ug = 0, adb = (unsigned)S.cl_off_box0;
    int st = n_oct0 > 0 ? kL0 : kNextWord, j0 = 0, w = 0, gi = 0, cbase = 0, wa = -n_always;
#pragma unroll 1
    for (;;) {
        if (st <= kL2) {
            const unsigned m8 = merge4(boxes(adb), boxes(adb + 3), boxes(adb + 6), boxes(adb + 9));
            nboxes += 8;
            if (st == kL0) {
                uw |= (unsigned long long)may_hit8(m8) << (56 - 8 * j0);
                j0++;
                if (j0 < n_oct0) { adb += 12; continue; }
                uw &= ~0ull << (64 - n_words);   // padding words (and "everything" of a far-out ray) stay inside the tables
                st = kNextWord;
            } else if (st == kL1) {
                ug = may_hit8(m8);
                cm = 0;
                st = kNextGroup;
            } else {
                cm |= (unsigned long long)(off ? 0xffu : (~m8 & 0xffu)) << (56 - 8 * gi);
                st = kNextGroup;
            }
        }
        if (st == kNextGroup) {
            if (ug) {
                const int gbit = 31 - __clz(ug);
                ug &= ~(1u << gbit);
                gi = 7 - gbit;
                adb = (unsigned)S.cl_off_box2 + (unsigned)(w * 8 + gi) * 12u;
                st = kL2;
                continue;
            }
            if (!has) cm = 0;
            const unsigned lo = __reduce_or_sync(kFull, (unsigned)cm), hi = __reduce_or_sync(kFull, (unsigned)(cm >> 32));
            ucm = ((unsigned long long)hi << 32) | lo;
            cbase = w * 64;
            st = ucm ? kChunks : kNextWord;
        }
        if (st == kNextWord) {
            if (wa < 0) {
                cbase = (S.cl_real_groups + (wa + n_always)) * 8;
                ucm = (unsigned long long)(wa == -1 ? S.cl_always_last : 0xffu) << 56;
                wa++;
                if (ucm == 0) continue;
            } else if (uw) {
                w = __clzll((long long)uw);
                uw &= ~(0x8000000000000000ull >> w);
                adb = (unsigned)S.cl_off_box1 + (unsigned)w * 12u;
                st = kL1;
                continue;
            } else break;
        }
        st = kNextWord;
        // ---- the marked chunks of this word through the pair pre-filter (software-pipelined by half chunks) ----
        int pos = __clzll((long long)ucm);
        unsigned ad = (unsigned)(cbase + pos) * 8u;
        float4 a0, a1, a2, a3, b0, b1, b2, b3;
        ldg4(a0, ad); ldg4(a1, ad + 1); ldg4(a2, ad + 2); ldg4(a3, ad + 3);
#pragma unroll 1
        for (;;) {
            ldg4(b0, ad + 4); ldg4(b1, ad + 5); ldg4(b2, ad + 6); ldg4(b3, ad + 7);
            const unsigned long long rest = ucm & ~(0x8000000000000000ull >> pos);
            const unsigned p0 = pair(a0, a1), p1 = pair(a2, a3);
            const int npos = rest ? __clzll((long long)rest) : pos;
            const unsigned nad = (unsigned)(cbase + npos) * 8u;
            ldg4(a0, nad); ldg4(a1, nad + 1); ldg4(a2, nad + 2); ldg4(a3, nad + 3);
            const unsigned mask = merge4(p0, p1, pair(b0, b1), pair(b2, b3));
            nchunks++;
            unsigned m = has ? (~mask & 0xffu) : 0u;  // 1 = must be tested exactly
            if (m) {
                const int chunk = cbase + pos;
                if (ncand > kCand - 8) {  // list about to overflow (rare): run the exact test on what is queued
                    T fx_ = ox, fy_ = oy, fz_ = oz, gx_ = dx, gy_ = dy, gz_ = dz;
                    if (parked) {
                        fx_ = T(parked[0]); fy_ = T(parked[TPB]); fz_ = T(parked[2 * TPB]);
                        gx_ = T(parked[3 * TPB]); gy_ = T(parked[4 * TPB]); gz_ = T(parked[5 * TPB]);
                    }
                    const BestHit<T> bh = flush_candidates_lex<T, FMA, TPB>(ggeo, cand, ncand, fx_, fy_, fz_, gx_, gy_, gz_, best_t, best);
                    best_t = bh.t; best = bh.id;
                    ncand = 0;
                }
                do {  // bit 7-u <-> slot chunk*8+u
                    const int bit = 31 - __clz(m);
                    if (TRAY_CHECK(ncand < kCand && chunk >= 0 && chunk * 8 + 7 - bit < (S.cl_off_box2 >> 3) * 8))
                        cand[ncand * TPB] = gids[chunk * 8 + 7 - bit];
                    ncand++;
                    m &= ~(1u << bit);
                } while (m);
            }
            if (!rest) break;
            ucm = rest; pos = npos; ad = nad;
        }
    }
}

// Exchange area of the regroup layout (one per CTA, shared memory): the path state of every lane, SoA.
template <int TPB>
struct RegroupBuf {
    double f[7][TPB];               // O, D, best_t
    unsigned long long r[2][TPB];   // generator state
    int i[5][TPB];                  // depth_left, sp, sample index, best (-1: miss, -2: no path), stack slot
    int count[5][TPB / 32];         // lanes per (class, warp)
};

// REGROUP (layout "per-material queues inside the megakernel"): after Scene.Hit the TPB paths of the CTA are sorted by what
// happens next -- miss (sky, finish, regenerate) | Lambertian | Metal | Dielectric | no path -- through shared memory, so
// that the divergent tail of the iteration (scatter, generators, unwind, regeneration) runs on warps that are mostly of
// one kind. Paths have no lane affinity; the attenuation stack lives in global memory under a slot id that travels with
// the path. Results are bit-identical to the plain layout (pure data movement).
// Register budget of the trace kernel: resident CTAs per SM (MINB), or -- tuning builds -- an explicit register count
// (TRAY_MAXNREG; e.g. 112 registers x 96 threads x 6 CTAs = 18 warps per SM, which no whole CTA count of 128 threads gives).
#ifdef TRAY_MAXNREG
#define TRAY_TRACE_BOUNDS(TPB, MINB) __maxnreg__(TRAY_MAXNREG)
#else
#define TRAY_TRACE_BOUNDS(TPB, MINB) __launch_bounds__(TPB, MINB)
#endif
template <typename T, bool FMA, int TPB, int MINB, int GEO, bool REGROUP = false, int BATCH = kBatch>
__global__ void TRAY_TRACE_BOUNDS(TPB, MINB) trace_kernel(const __grid_constant__ TraceArgs A, const __grid_constant__ DevScene<T> S,
                                                          const __grid_constant__ GeoArg<T, GEO> GP) {
    typedef typename Vec4T<T>::type T4;
    constexpr int CH = TRAY_CH;  // spheres per candidate-mask chunk (n_pad is a multiple of 8)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T4* sgeo = reinterpret_cast<T4*>(smem_raw);
    const size_t geo_bytes = GEO == kGeoShared ? (size_t)S.n_pad * sizeof(T4)
                           : (GEO == kGeoFilter ? (size_t)S.n_pad * 16 : (GEO == kGeoCluster ? (size_t)S.cl_blob_f4 * 16 : 0));
    float4* sfp = reinterpret_cast<float4*>(smem_raw);  // kGeoFilter: n_pad/2 pairs x 2 float4; kGeoCluster: the cluster blob
    ZigTables* zig = reinterpret_cast<ZigTables*>(smem_raw + geo_bytes);
    uint16_t* cand_all = reinterpret_cast<uint16_t*>(smem_raw + geo_bytes + sizeof(ZigTables));
    const size_t pool_off = (geo_bytes + sizeof(ZigTables) + (size_t)kCand * TPB * sizeof(uint16_t) + 15) & ~(size_t)15;
    CamRay* gpool = reinterpret_cast<CamRay*>(smem_raw + pool_off) + (threadIdx.x >> 5) * kGenSlots;  // this warp's generated camera rays
    RegroupBuf<TPB>* xb = reinterpret_cast<RegroupBuf<TPB>*>(smem_raw + pool_off + kGenPoolBytes<TPB>);
    const int tid = threadIdx.x;
    if (GEO == kGeoShared)
        for (int i = tid; i < S.n_pad; i += TPB) sgeo[i] = S.geo[i];
    if (GEO == kGeoFilter)
        for (int i = tid; i < S.n_pad; i += TPB) sfp[i] = S.fpair[i];
    if (GEO == kGeoCluster)
        for (int i = tid; i < S.cl_blob_f4; i += TPB) sfp[i] = S.cl_blob[i];
    zig_load(zig, tid, TPB);
    __syncthreads();
    const T4* __restrict__ ggeo = S.geo;
    uint16_t* cand = cand_all + tid;  // this lane's list: cand[k*TPB]

    const int lane = tid & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    bool has = false, exhausted = false;
    unsigned pool_next = 0, pool_end = 0, my_li = 0;  // local sample indices of this pass (< 2^32 by construction)
    unsigned gen_first = 0, gen_avail = 0, gen_base = 0;  // TRAY_GEN_INKERNEL: generated rays waiting in this warp's pool
    (void)gen_first; (void)gen_avail; (void)gen_base; (void)gpool;
    const unsigned n_samples = (unsigned)A.n_samples;
    // Idle lanes carry a ray that certainly misses everything (every sphere is far behind it), so the
    // hot loop needs no "has a path" test.
    V3<T> O = mk<T>(T(0), T(1e18), T(0)), D = mk<T>(T(0), T(1), T(0));
    Pcg rng = pcg_new_idx(0, 0);
    int depth_left = 0, sp = 0;
    int skip = -1;  // fp32 fast path (TRAY_FP32_SKIP_ORIGIN): sphere the current ray leaves outwards, -1 = none
    (void)skip;
    constexpr bool kForward = sizeof(T) == 4 && TRAY_FP32_FORWARD && !REGROUP;  // fp32 fast path: throughput carried forward, no id stack
    V3<T> thr = mk<T>(T(1), T(1), T(1));
    (void)thr;
    uint16_t stk[(REGROUP || kForward) ? 1 : kMaxDepth];
    unsigned slot = blockIdx.x * TPB + tid;  // regroup layout: where this path's attenuation stack lives
    unsigned long long nseg = 0, ntests = 0, nbox = 0;
    unsigned nexh = 0, ndone = 0, ntests_blk = 0, nbox_blk = 0;
    const int n_pad = S.n_pad;

    for (;;) {
        // ---------------- path regeneration (lane level) ----------------
        if (!exhausted) {
            unsigned need = __ballot_sync(kFull, !has);
            while (need) {
#if TRAY_GEN_INKERNEL
                if (gen_avail == 0 && pool_next >= pool_end) {
#else
                if (pool_next >= pool_end) {
#endif
                    unsigned long long base = 0;
                    if (lane == 0) base = atomicAdd(A.counter, (unsigned long long)BATCH);
                    base = __shfl_sync(kFull, base, 0);
                    if (base >= A.n_samples) { exhausted = true; break; }
                    pool_next = (unsigned)base;
                    pool_end = (unsigned)base + BATCH < n_samples ? (unsigned)base + BATCH : n_samples;
                }
#if TRAY_GEN_INKERNEL
                if (gen_avail == 0) {  // the whole warp generates the camera rays of the next (up to) kGenSlots samples of its batch
                    const unsigned n = pool_end - pool_next < kGenSlots ? pool_end - pool_next : kGenSlots;
                    if ((unsigned)lane < n) {
                        Pcg g;
                        V3<double> o64, d64;
                        camera_sample(A, pool_next + lane, g, o64, d64);
                        double2* q = reinterpret_cast<double2*>(gpool + lane);
                        q[0] = make_double2(o64.x, o64.y); q[1] = make_double2(o64.z, d64.x); q[2] = make_double2(d64.y, d64.z);
                        reinterpret_cast<ulonglong2*>(q)[3] = make_ulonglong2(g.hi, g.lo);
                    }
                    __syncwarp();
                    gen_first = 0; gen_avail = n; gen_base = pool_next;
                    pool_next += n;
                }
                unsigned avail = gen_avail;
                unsigned rank = __popc(need & lt_mask);
                if (!has && rank < avail && TRAY_CHECK(gen_first + rank < kGenSlots)) {
                    my_li = gen_base + gen_first + rank;
                    const double2* q = reinterpret_cast<const double2*>(gpool + gen_first + rank);
#else
                unsigned avail = pool_end - pool_next;
                unsigned rank = __popc(need & lt_mask);
                if (!has && rank < avail) {
                    my_li = pool_next + rank;
                    // the camera ray and the generator state after its draws were made ahead by camera_ray_kernel
                    const double2* q = reinterpret_cast<const double2*>(A.gen + my_li);
#endif
                    const double2 q0 = q[0], q1 = q[1], q2 = q[2];
                    const ulonglong2 q3 = reinterpret_cast<const ulonglong2*>(q)[3];
                    const V3<double> o64 = mk<double>(q0.x, q0.y, q1.x), d64 = mk<double>(q1.y, q2.x, q2.y);
                    rng.hi = q3.x; rng.lo = q3.y;
                    O = mk<T>(T(o64.x), T(o64.y), T(o64.z));
                    D = mk<T>(T(d64.x), T(d64.y), T(d64.z));
                    depth_left = A.max_depth;
                    sp = 0;
                    skip = -1;
                    if constexpr (kForward) thr = mk<T>(T(1), T(1), T(1));
                    has = true;
                }
                unsigned cnt = __popc(need);
#if TRAY_GEN_INKERNEL
                __syncwarp();
                { unsigned took = cnt < avail ? cnt : avail; gen_first += took; gen_avail -= took; }
#else
                pool_next += cnt < avail ? cnt : avail;
#endif
                need = __ballot_sync(kFull, !has);
            }
        }
        const bool warp_alive = __any_sync(kFull, has);
        if constexpr (!REGROUP) {
            if (!warp_alive) break;
        }  // regroup layout: the CTA leaves together, decided at the exchange barrier below

        // ---------------- Scene.Hit: brute force over all spheres ----------------
        T best_t = t_inf<T>();
        int best = -1;
        V3<T> ud = mk<T>(T(0), T(1), T(0));
        if (warp_alive) {  // (regroup layout: warps left without a path near the end of a pass skip the scan)
#if TRAY_PARK_STATE
        // Park what the scan does not need (fp64 ray, generator, counters of the path) in shared memory, so that the loop's
        // registers are not shared with ~20 registers of path state.
        __shared__ double s_park[11][TPB];
        __shared__ int s_parki[3][TPB];
        volatile double* pk = &s_park[0][tid];
        volatile int* pki = &s_parki[0][tid];
        constexpr bool kPark = (GEO == kGeoFilter || GEO == kGeoCluster || GEO == kGeoClusterBig) && !REGROUP && sizeof(T) == 8;  // (fp32 fast path: state stays in registers -- 72 registers x 28 warps measured best, 33.6 vs 36.2 ms parked)
        if constexpr (kPark) {
            pk[0] = (double)O.x; pk[TPB] = (double)O.y; pk[2 * TPB] = (double)O.z;
            pk[3 * TPB] = (double)D.x; pk[4 * TPB] = (double)D.y; pk[5 * TPB] = (double)D.z;
            pk[6 * TPB] = __longlong_as_double((long long)rng.hi); pk[7 * TPB] = __longlong_as_double((long long)rng.lo);
            pki[0] = depth_left; pki[TPB] = sp; pki[2 * TPB] = (int)my_li;
        }
#endif
        // Unit(r.Direction) is needed by the sky (objects.go:69), Metal (materials.go:29) and Dielectric (materials.go:52): computed
        // once for every lane, and BEFORE the scan, whose fp32 filters take their normalised direction from it.
        { const T l = tsqrt_hot(len2(D)); ud = div3_hot(D, l); }  // Unit(r.Direction), ray/vec3.go
#if TRAY_PARK_STATE
        if constexpr (kPark) { pk[8 * TPB] = (double)ud.x; pk[9 * TPB] = (double)ud.y; pk[10 * TPB] = (double)ud.z; }
#endif
        T ox = O.x, oy = O.y, oz = O.z, dx = D.x, dy = D.y, dz = D.z;
        T a = len2(D);  // LengthSquared(r.Direction), objects.go:83 (same bits for every sphere)
        int ncand = 0;
        unsigned mask_prev = 0;
        if constexpr (GEO == kGeoBVH) {
            bvh_closest_hit<T, FMA>(S, ggeo, has, ox, oy, oz, dx, dy, dz, a, best_t, best, ntests_blk);
            if (ntests_blk > 0x40000000u) { ntests += ntests_blk; ntests_blk = 0; }
        } else if constexpr (GEO == kGeoFilter || GEO == kGeoCluster || GEO == kGeoClusterBig) {
            unsigned nchunks = 0, nboxes = 0;
            (void)nchunks; (void)nboxes;
#if TRAY_PARK_STATE
            if constexpr (kPark) {
                if constexpr (GEO == kGeoCluster) cluster_scan<T, FMA, TPB>(S, sfp, ggeo, cand, has, ox, oy, oz, dx, dy, dz, a, ud.x, ud.y, ud.z, best_t, best, ncand, nchunks, nboxes, pk);
                else if constexpr (GEO == kGeoClusterBig) cluster_scan_big<T, FMA, TPB>(S, ggeo, cand, has, ox, oy, oz, dx, dy, dz, a, ud.x, ud.y, ud.z, best_t, best, ncand, nchunks, nboxes, pk);
                else
                filter_scan<T, FMA, TPB>(S, sfp, ggeo, cand, has, ox, oy, oz, dx, dy, dz, a, best_t, best, ncand, mask_prev, pk);
                ox = T(pk[0]); oy = T(pk[TPB]); oz = T(pk[2 * TPB]); dx = T(pk[3 * TPB]); dy = T(pk[4 * TPB]); dz = T(pk[5 * TPB]);
                a = (dx * dx + dy * dy) + dz * dz;
                O = mk<T>(ox, oy, oz); D = mk<T>(dx, dy, dz);
                rng.hi = (uint64_t)__double_as_longlong(pk[6 * TPB]); rng.lo = (uint64_t)__double_as_longlong(pk[7 * TPB]);
                depth_left = pki[0]; sp = pki[TPB]; my_li = (unsigned)pki[2 * TPB];
                ud = mk<T>(T(pk[8 * TPB]), T(pk[9 * TPB]), T(pk[10 * TPB]));
            } else
#endif
            {
                if constexpr (GEO == kGeoCluster) cluster_scan<T, FMA, TPB>(S, sfp, ggeo, cand, has, ox, oy, oz, dx, dy, dz, a, ud.x, ud.y, ud.z, best_t, best, ncand, nchunks, nboxes);
                else if constexpr (GEO == kGeoClusterBig) cluster_scan_big<T, FMA, TPB>(S, ggeo, cand, has, ox, oy, oz, dx, dy, dz, a, ud.x, ud.y, ud.z, best_t, best, ncand, nchunks, nboxes);
                else filter_scan<T, FMA, TPB>(S, sfp, ggeo, cand, has, ox, oy, oz, dx, dy, dz, a, best_t, best, ncand, mask_prev);
            }
            if constexpr (GEO == kGeoCluster || GEO == kGeoClusterBig) {
                if (has) { ntests_blk += 8u * nchunks; nbox_blk += nboxes; }
                if (ntests_blk > 0x40000000u) { ntests += ntests_blk; ntests_blk = 0; }
                if (nbox_blk > 0x40000000u) { nbox += nbox_blk; nbox_blk = 0; }
            }
        } else {
        // Branch-free body: every test contributes one bit ("may be hit") to the chunk mask through a funnel
        // shift. The mask of chunk k is examined while chunk k+1 is in flight, so no branch waits on FP64 results.
#pragma unroll 1
        for (int i = 0; i < n_pad; i += CH) {
            unsigned mask = 0;
#pragma unroll
            for (int u = 0; u < CH; u++) {
                T4 g;
                if constexpr (GEO == kGeoParam) g = GP.g[i + u];
                else if constexpr (GEO == kGeoShared) g = sgeo[i + u];
                else g = ggeo[i + u];
                T h, c, disc;
                sphere_terms<T, FMA>(ox, oy, oz, dx, dy, dz, a, g.x, g.y, g.z, g.w, h, c, disc);
                mask = __funnelshift_l((unsigned)miss_bits(h, c, disc), mask, 1);  // shifts in the "missed" sign bit
            }
            mask = ~mask & ((1u << CH) - 1u);  // 1 = may be hit
            if (mask_prev)
                push_candidates<T, FMA, TPB, CH>(ggeo, cand, &ncand, mask_prev, i - CH, ox, oy, oz, dx, dy, dz, a, &best_t, &best);
            mask_prev = mask;
        }
        }
        if constexpr (GEO == kGeoCluster || GEO == kGeoClusterBig) {
            resolve_candidates_lex<T, FMA, TPB>(ggeo, cand, ncand, ox, oy, oz, dx, dy, dz, a, best_t, best, skip);
        } else {
            if (mask_prev)
                push_candidates<T, FMA, TPB, CH>(ggeo, cand, &ncand, mask_prev, n_pad - CH, ox, oy, oz, dx, dy, dz, a, &best_t, &best);
            resolve_candidates<T, FMA, TPB>(ggeo, cand, ncand, ox, oy, oz, dx, dy, dz, a, best_t, best);
        }
        }

        // ---------------- regroup: sort the CTA's paths by what happens next ----------------
        if constexpr (REGROUP) {
            const int warp = tid >> 5;
            int cls = 4;  // no path
            if (has) cls = best < 0 ? 0 : 1 + (int)S.kind[best];
            unsigned bal[5];
#pragma unroll
            for (int c = 0; c < 5; c++) bal[c] = __ballot_sync(kFull, cls == c);
            unsigned mine = bal[0];
#pragma unroll
            for (int c = 1; c < 5; c++) if (cls == c) mine = bal[c];
            const int rank = __popc(mine & lt_mask);
            if (lane < 5) {
                unsigned v = bal[0];
#pragma unroll
                for (int c = 1; c < 5; c++) if (lane == c) v = bal[c];
                xb->count[lane][warp] = __popc(v);
            }
            if (!__syncthreads_or(has ? 1 : 0)) break;  // no path left in the CTA (and none to be had: regeneration ran dry)
            int dst = rank;
#pragma unroll
            for (int c = 0; c < 5; c++)
#pragma unroll
                for (int w = 0; w < TPB / 32; w++) {
                    const int n = xb->count[c][w];
                    if (c < cls || (c == cls && w < warp)) dst += n;
                }
            if (!TRAY_CHECK(dst >= 0 && dst < TPB)) dst = tid;
            xb->f[0][dst] = (double)O.x; xb->f[1][dst] = (double)O.y; xb->f[2][dst] = (double)O.z;
            xb->f[3][dst] = (double)D.x; xb->f[4][dst] = (double)D.y; xb->f[5][dst] = (double)D.z;
            xb->f[6][dst] = (double)best_t;
            xb->r[0][dst] = rng.hi; xb->r[1][dst] = rng.lo;
            xb->i[0][dst] = depth_left; xb->i[1][dst] = sp; xb->i[2][dst] = (int)my_li;
            xb->i[3][dst] = has ? best : -2; xb->i[4][dst] = (int)slot;
            __syncthreads();
            O = mk<T>(T(xb->f[0][tid]), T(xb->f[1][tid]), T(xb->f[2][tid]));
            D = mk<T>(T(xb->f[3][tid]), T(xb->f[4][tid]), T(xb->f[5][tid]));
            best_t = T(xb->f[6][tid]);
            rng.hi = xb->r[0][tid]; rng.lo = xb->r[1][tid];
            depth_left = xb->i[0][tid]; sp = xb->i[1][tid]; my_li = (unsigned)xb->i[2][tid];
            best = xb->i[3][tid]; slot = (unsigned)xb->i[4][tid];
            has = best != -2;
        }

        // ---------------- RayColor step (ray/objects.go:49-62) ----------------
        if (has) {
            nseg++;
            bool finish = false;
            V3<T> col = mk<T>(T(0), T(0), T(0));
            if constexpr (REGROUP) ud = unit(D);  // (the exchange moved the path to another lane; Unit(D) is not part of the exchanged state)
            if (best < 0) {
                T ab = T(0.5) * (ud.y + T(1));  // AmbientLight.Hit, objects.go:68-73
                col = mk<T>(T(S.bg_a[0]), T(S.bg_a[1]), T(S.bg_a[2])) * (T(1) - ab) + mk<T>(T(S.bg_b[0]), T(S.bg_b[1]), T(S.bg_b[2])) * ab;
                finish = true;
            } else {
                T4 g = ggeo[best];
                V3<T> P, N;
                bool front;
                hit_record<T>(O, D, best_t, mk<T>(g.x, g.y, g.z), S.radius[best], P, N, front);
                const int kind = S.kind[best];
                const double4 prm = S.params[best];
                // one shared UnitVector draw site for Lambertian (always) and Metal (iff Fuzz > 0)
                const bool need_uv = kind == 0 || (kind == 1 && prm.w > 0.0);
                V3<T> uv = mk<T>(T(0), T(0), T(0));
                if (need_uv) {
                    if constexpr (sizeof(T) == 4 && TRAY_FP32_RNG) uv = pcg_unit_vector_f32(rng, zig, S.unitvec_variant);  // fp32 fast path: float normals
                    else { V3<double> u64 = pcg_unit_vector(rng, zig, S.unitvec_variant); uv = mk<T>(T(u64.x), T(u64.y), T(u64.z)); }
                }
                bool scattered = true;
                V3<T> D2;
                if (kind == 0) {  // Lambertian.Scatter, materials.go:13-21
                    D2 = N + uv;
                    if (near_zero(D2)) D2 = N;
                } else if (kind == 1) {  // Metal.Scatter, materials.go:28-38
                    D2 = reflect(ud, N);
                    if (prm.w > 0.0) D2 = D2 + uv * T(prm.w);
                    scattered = dot(D2, N) > T(0);
                } else {  // Dielectric.Scatter, materials.go:44-64
                    T ri = T(prm.x);
                    T ratio = front ? tdiv(T(1), ri) : ri;
                    T cosTheta = tmin2(dot(vneg(ud), N), T(1));
                    T sinTheta = tsqrt(T(1) - cosTheta * cosTheta);
                    bool refl = ratio * sinTheta > T(1);
                    if (!refl) refl = (double)reflectance(cosTheta, ratio) > pcg_f64(rng);  // drawn only when it can refract
                    D2 = refl ? reflect(ud, N) : refract(ud, N, ratio);
                }
                if (!scattered) {
                    finish = true;  // absorbed: black
                } else {
                    if (kind != 2) {  // attenuation = albedo; Dielectric's (1,1,1) is an exact identity
                        if constexpr (kForward) thr = vmul(thr, mk<T>(T(prm.x), T(prm.y), T(prm.z)));
                        else if constexpr (REGROUP) A.stk_g[(size_t)sp * A.n_slots + slot] = (uint16_t)best;
                        else if (TRAY_CHECK(sp >= 0 && sp < kMaxDepth && best >= 0 && best < S.n)) stk[sp] = (uint16_t)best;
                        sp++;
                    }
                    if constexpr (sizeof(T) == 4 && TRAY_FP32_SKIP_ORIGIN && !REGROUP) skip = ((dot(D2, N) > T(0)) == front) ? best : -1;  // (not part of the regroup exchange)
                    O = P; D = D2;
                    depth_left--;
                    if (depth_left <= 0) { finish = true; nexh++; }  // RayColor(depth<=0) = black
                }
            }
            if (finish) {
                if constexpr (kForward) col = vmul(thr, col);
                else if (col.x != T(0) || col.y != T(0) || col.z != T(0)) {
#pragma unroll 1
                    for (int k = sp - 1; k >= 0; k--) {  // Mul(attenuation, ...) applied on unwind, objects.go:56 (one copy: instruction-cache footprint)
                        int id;
                        if constexpr (REGROUP) id = A.stk_g[(size_t)k * A.n_slots + slot];
                        else id = stk[k];
                        double4 prm = S.params[id];
                        col = vmul(mk<T>(T(prm.x), T(prm.y), T(prm.z)), col);
                    }
                }
                double* out = A.scratch + 3ull * my_li;
                if (TRAY_CHECK((unsigned long long)my_li < A.n_samples)) { out[0] = (double)col.x; out[1] = (double)col.y; out[2] = (double)col.z; }
                has = false;
                O = mk<T>(T(0), T(1e18), T(0)); D = mk<T>(T(0), T(1), T(0));
                ndone++;
            }
        }
    }

    // stats: warp-reduce, one atomic per warp
    ntests += ntests_blk;
    nbox += nbox_blk;
    for (int off = 16; off > 0; off >>= 1) {
        ntests += __shfl_down_sync(kFull, ntests, off);
        nbox += __shfl_down_sync(kFull, nbox, off);
        nseg += __shfl_down_sync(kFull, nseg, off);
        nexh += __shfl_down_sync(kFull, nexh, off);
        ndone += __shfl_down_sync(kFull, ndone, off);
    }
    if (lane == 0) {
        atomicAdd(&A.stats[0], nseg);
        atomicAdd(&A.stats[1], (unsigned long long)nexh);
        if (ntests) atomicAdd(&A.stats[3], ntests);
        if (nbox) atomicAdd(&A.stats[4], nbox);

        if (A.progress) atomicAdd(A.progress, (unsigned long long)ndone);
    }
}

// ---------------------------------------------------------------------------------------------
struct ResolveArgs {
    const double* scratch;
    unsigned long long n_pixels;      // pixels in this pass
    unsigned long long pass_pixel0;   // first local pixel
    int spp_local;
    double inv_spp;                   // 1.0 / float64(NumRaysPerPixel), ray/tracer.go:123
    int partial;                      // 1: store the raw partial sum only (sample split / subsets); 2: continue the stored sum
    unsigned char* rgba;              // local image, 4 B/pixel
    double* hdr;                      // local, 3 doubles/pixel (mean, or partial sum when partial)
    const double* srgb_thr;
};

__global__ void __launch_bounds__(256) resolve_kernel(const ResolveArgs R) {
    // one thread per (pixel, channel): three times the loads in flight of a thread per pixel, neighbouring threads read
    // neighbouring doubles (the kernel is HBM-bound: 24 B per sample read once)
    const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long p = t / 3ull;
    const int c = (int)(t - p * 3ull);
    if (p >= R.n_pixels) return;
    const double* s = R.scratch + 3ull * p * (unsigned)R.spp_local + c;
    const unsigned long long lp = R.pass_pixel0 + p;
    double sum = 0.0;
    if (R.partial == 2) sum = R.hdr[3 * lp + c];  // progressive: same running sum
#pragma unroll 8
    for (int j = 0; j < R.spp_local; j++) sum = sum + s[3 * j];  // colorSum = Add(colorSum, color), sample order (ray/tracer.go:143)
    if (R.partial) { R.hdr[3 * lp + c] = sum; return; }
    const double v = sum * R.inv_spp;  // SMul(colorSum, colorSumDiv)
    if (R.hdr) R.hdr[3 * lp + c] = v;
    R.rgba[4 * lp + c] = linear_to_srgb(R.srgb_thr, v);
    if (c == 0) R.rgba[4 * lp + 3] = 255;
}

// MaxDepth == 0: RayColor returns black for every sample (ray/objects.go:50-52), the pixel is (0,0,0,255).
__global__ void __launch_bounds__(256) black_kernel(unsigned char* rgba, double* hdr, unsigned long long n_pixels) {
    unsigned long long p = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pixels) return;
    reinterpret_cast<uchar4*>(rgba)[p] = make_uchar4(0, 0, 0, 255);
    hdr[3 * p] = 0.0; hdr[3 * p + 1] = 0.0; hdr[3 * p + 2] = 0.0;
}

// Sample-split epilogue: sum the per-device partial sums (peer pointers over NVLink, device order),
// scale, convert and store -- the reduction and the sRGB store are one kernel.
struct CombineArgs {
    const double* partial[8];
    int n_parts;
    unsigned long long n_pixels;
    double inv_spp;
    unsigned char* rgba;
    double* hdr;
    const double* srgb_thr;
};

__global__ void __launch_bounds__(256) combine_kernel(const CombineArgs C) {
    unsigned long long p = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= C.n_pixels) return;
    double sx = 0.0, sy = 0.0, sz = 0.0;
    for (int g = 0; g < C.n_parts; g++) {
        const double* q = C.partial[g] + 3 * p;
        sx = sx + q[0]; sy = sy + q[1]; sz = sz + q[2];
    }
    double cx = sx * C.inv_spp, cy = sy * C.inv_spp, cz = sz * C.inv_spp;
    if (C.hdr) { C.hdr[3 * p] = cx; C.hdr[3 * p + 1] = cy; C.hdr[3 * p + 2] = cz; }
    uchar4 px;
    px.x = linear_to_srgb(C.srgb_thr, cx);
    px.y = linear_to_srgb(C.srgb_thr, cy);
    px.z = linear_to_srgb(C.srgb_thr, cz);
    px.w = 255;
    reinterpret_cast<uchar4*>(C.rgba)[p] = px;
}

// ---------------------------------------------------------------------------------------------
// Conformance mode: the reference's own stream semantics. One warp per chunk; all 32 lanes carry the
// same path state and RNG (uniform, redundant), and split the sphere loop 32 ways.
// ---------------------------------------------------------------------------------------------
struct RefArgs {
    DevCamera cam;
    int width, spp, max_depth;
    double ray_radius;
    unsigned long long seed;
    int row_begin, row_end;  // rows covered by all chunks
    int chunk_rows, n_chunks;
    long long idx_override;  // >= 0: the single chunk uses this stream index (RenderLines(idx,..)); else idx = chunk start row
    unsigned char* rgba;     // local image covering rows [row_begin,row_end)
    double* hdr;
    const double* srgb_thr;
    unsigned long long* stats;
};

template <bool FMA>
__global__ void __launch_bounds__(32) reference_stream_kernel(const RefArgs A, const DevScene<double> S) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double4* geo = reinterpret_cast<double4*>(smem_raw);
    ZigTables* zig = reinterpret_cast<ZigTables*>(smem_raw + (size_t)S.n_pad * sizeof(double4));
    const int lane = threadIdx.x;
    for (int i = lane; i < S.n_pad; i += 32) geo[i] = S.geo[i];
    zig_load(zig, lane, 32);
    __syncwarp();
    const int chunk = blockIdx.x;
    if (chunk >= A.n_chunks) return;
    const int ys = A.row_begin + chunk * A.chunk_rows;
    const int ye = ys + A.chunk_rows < A.row_end ? ys + A.chunk_rows : A.row_end;
    Pcg rng = pcg_new_idx(A.idx_override >= 0 ? (unsigned long long)A.idx_override : (unsigned long long)(long long)ys, A.seed);
    const double inv_spp = 1.0 / (double)A.spp;
    uint16_t stk[kMaxDepth];
    unsigned long long nseg = 0, nexh = 0;
    const double inf = t_inf<double>();

    for (int y = ys; y < ye; y++) {
        for (int x = 0; x < A.width; x++) {
            V3<double> sum = mk<double>(0.0, 0.0, 0.0);
            for (int s = 0; s < A.spp; s++) {
                double jx = 0.0, jy = 0.0;
                if (A.spp > 1) pcg_in_disc(rng, A.ray_radius, jx, jy, A.cam.indisc_variant);
                V3<double> O, D;
                get_ray(A.cam, rng, (double)x, (double)y, jx, jy, O, D);
                int depth_left = A.max_depth, sp = 0;
                V3<double> col = mk<double>(0.0, 0.0, 0.0);
                for (;;) {
                    // cooperative Scene.Hit: lane l scans spheres l, l+32, ... with the running-closest rule,
                    // then a lexicographic (t, id) warp minimum restores "ties -> lowest index".
                    const double a = len2(D);
                    double bt = inf;
                    int b = -1;
                    for (int i = lane; i < S.n; i += 32) {
                        double4 g = geo[i];
                        double h, c, disc, root;
                        sphere_terms<double, FMA>(O.x, O.y, O.z, D.x, D.y, D.z, a, g.x, g.y, g.z, g.w, h, c, disc);
                        if (sphere_root<double>(h, a, disc, 1e-6, bt, root)) { bt = root; b = i; }
                    }
                    for (int off = 16; off > 0; off >>= 1) {
                        double ot = __shfl_xor_sync(kFull, bt, off);
                        int ob = __shfl_xor_sync(kFull, b, off);
                        if (ob >= 0 && (b < 0 || ot < bt || (ot == bt && ob < b))) { bt = ot; b = ob; }
                    }
                    nseg++;
                    if (b < 0) { col = background<double>(S.bg_a, S.bg_b, D); break; }
                    double4 g = geo[b];
                    V3<double> P, N, O2, D2;
                    bool front, alb;
                    hit_record<double>(O, D, bt, mk<double>(g.x, g.y, g.z), S.radius[b], P, N, front);
                    bool scattered = scatter<double>(S.kind[b], S.params[b], rng, zig, D, P, N, front, O2, D2, alb, S.unitvec_variant);
                    if (!scattered) break;
                    if (alb) stk[sp++] = (uint16_t)b;
                    O = O2; D = D2;
                    if (--depth_left <= 0) { nexh++; break; }
                }
                if (col.x != 0.0 || col.y != 0.0 || col.z != 0.0)
                    for (int k = sp - 1; k >= 0; k--) {
                        double4 prm = S.params[stk[k]];
                        col = vmul(mk<double>(prm.x, prm.y, prm.z), col);
                    }
                sum = sum + col;
            }
            if (lane == 0) {
                V3<double> c = sum * inv_spp;
                size_t lp = (size_t)(y - A.row_begin) * A.width + x;
                if (A.hdr) { A.hdr[3 * lp] = c.x; A.hdr[3 * lp + 1] = c.y; A.hdr[3 * lp + 2] = c.z; }
                uchar4 px;
                px.x = linear_to_srgb(A.srgb_thr, c.x);
                px.y = linear_to_srgb(A.srgb_thr, c.y);
                px.z = linear_to_srgb(A.srgb_thr, c.z);
                px.w = 255;
                reinterpret_cast<uchar4*>(A.rgba)[lp] = px;
            }
        }
    }
    if (lane == 0) { atomicAdd(&A.stats[0], nseg); atomicAdd(&A.stats[1], nexh); }
}

// ---------------------------------------------------------------------------------------------
// RNG-free parity probe: Scene.Hit for the pixel-centre pinhole primary ray of every pixel.
// ---------------------------------------------------------------------------------------------
template <typename T, bool FMA>
__global__ void __launch_bounds__(128) first_hit_kernel(DevCamera cam, DevScene<T> S, int width, int height,
                                                        int* id, double* t, double* normal, unsigned char* front) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= width * height) return;
    int y = p / width, x = p - y * width;
    cam.aperture = 0.0;
    Pcg dummy = pcg_new_idx(0, 0);
    V3<double> o64, d64;
    get_ray(cam, dummy, (double)x, (double)y, 0.0, 0.0, o64, d64);
    V3<T> O = mk<T>(T(o64.x), T(o64.y), T(o64.z)), D = mk<T>(T(d64.x), T(d64.y), T(d64.z));
    T a = len2(D), bt = t_inf<T>();
    int b = -1;
    for (int i = 0; i < S.n; i++) {
        typename Vec4T<T>::type g = S.geo[i];
        T h, c, disc, root;
        sphere_terms<T, FMA>(O.x, O.y, O.z, D.x, D.y, D.z, a, g.x, g.y, g.z, g.w, h, c, disc);
        if (sphere_root<T>(h, a, disc, front_epsilon<T>(), bt, root)) { bt = root; b = i; }
    }
    id[p] = b;
    t[p] = (double)bt;
    if (b >= 0) {
        typename Vec4T<T>::type g = S.geo[b];
        V3<T> P, N;
        bool ff;
        hit_record<T>(O, D, bt, mk<T>(g.x, g.y, g.z), S.radius[b], P, N, ff);
        normal[3 * p] = (double)N.x; normal[3 * p + 1] = (double)N.y; normal[3 * p + 2] = (double)N.z;
        front[p] = ff ? 1 : 0;
    } else {
        normal[3 * p] = normal[3 * p + 1] = normal[3 * p + 2] = 0.0;
        front[p] = 0;
    }
}

// ---------------------------------------------------------------------------------------------
// "Next" row 8(f)-1: what tray does with the image after Render (main.go:119-130), kept on the device:
// draw.BiLinear.Scale (x/image/draw Kernel scaler, separable triangle filter widened by the scale factor,
// float64 intermediates, 16-bit premultiplied Over onto a fresh image) and the half-block truecolor frame.
// ---------------------------------------------------------------------------------------------
struct BlSrc { int i, j; double inv, inv_ffff, center, arg; };

__device__ __forceinline__ double bl_weight(const BlSrc& s, int c) {
    double t = fabs((s.center - (double)c) * s.arg);
    return t >= 1.0 ? 0.0 : 1.0 - t;
}
__device__ __forceinline__ BlSrc bl_source(int x, int dw, int sw) {
    double scale = (double)sw / (double)dw;
    double half = 1.0, arg = 1.0;
    if (scale > 1) { half *= scale; arg = 1 / scale; }
    BlSrc s;
    s.center = ((double)x + 0.5) * scale - 0.5;
    s.arg = arg;
    int i = (int)floor(s.center - half);
    if (i < 0) i = 0;
    int j = (int)ceil(s.center + half);
    if (j > sw) { j = sw; if (j < i) j = i; }
    s.i = i; s.j = j;
    double total = 0.0;
    for (int c = i; c < j; c++) total += bl_weight(s, c);  // zero weights add nothing, like the reference's `continue`
    s.inv = 1 / total;
    s.inv_ffff = s.inv / 65535.0;
    return s;
}
__device__ __forceinline__ unsigned bl_ftou(double f) {
    int i = (int)(65535.0 * f + 0.5);
    return i > 0xffff ? 0xffffu : (i > 0 ? (unsigned)i : 0u);
}

__global__ void __launch_bounds__(128) scale_x_kernel(const uchar4* __restrict__ src, int sw, int sh, int dw, double4* __restrict__ tmp) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= dw || y >= sh) return;
    BlSrc s = bl_source(x, dw, sw);
    double p0 = 0, p1 = 0, p2 = 0, p3 = 0;
    for (int c = s.i; c < s.j; c++) {
        double w = bl_weight(s, c);
        if (w == 0) continue;
        uchar4 px = src[(size_t)y * sw + c];
        p0 += (double)((unsigned)px.x * 0x101u) * w;
        p1 += (double)((unsigned)px.y * 0x101u) * w;
        p2 += (double)((unsigned)px.z * 0x101u) * w;
        p3 += (double)((unsigned)px.w * 0x101u) * w;
    }
    tmp[(size_t)y * dw + x] = make_double4(p0 * s.inv_ffff, p1 * s.inv_ffff, p2 * s.inv_ffff, p3 * s.inv_ffff);
}

__global__ void __launch_bounds__(128) scale_y_kernel(const double4* __restrict__ tmp, int dw, int sh, int dh, uchar4* __restrict__ dst) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= dw || y >= dh) return;
    BlSrc s = bl_source(y, dh, sh);
    double p0 = 0, p1 = 0, p2 = 0, p3 = 0;
    for (int c = s.i; c < s.j; c++) {
        double w = bl_weight(s, c);
        if (w == 0) continue;
        double4 t = tmp[(size_t)c * dw + x];
        p0 += t.x * w; p1 += t.y * w; p2 += t.z * w; p3 += t.w * w;
    }
    if (p0 > p3) p0 = p3;
    if (p1 > p3) p1 = p3;
    if (p2 > p3) p2 = p3;
    unsigned q0 = bl_ftou(p0 * s.inv), q1 = bl_ftou(p1 * s.inv), q2 = bl_ftou(p2 * s.inv), q3 = bl_ftou(p3 * s.inv);
    // Over onto a freshly allocated (zero) destination: dst*pa1/0xffff contributes 0
    dst[(size_t)y * dw + x] = make_uchar4((unsigned char)(q0 >> 8), (unsigned char)(q1 >> 8), (unsigned char)(q2 >> 8), (unsigned char)(q3 >> 8));
}

// draw.NearestNeighbor.Scale with draw.Over onto a fresh image (tray -s < 1, main.go:124-125): x/image/draw samples the source
// pixel ((2*dx+1)*sw/(2*dw), (2*dy+1)*sh/(2*dh)) in integer arithmetic; the 16-bit premultiplied Over of an opaque
// pixel onto zero gives (v*0x101) >> 8 = v, for a translucent one (v*0x101) >> 8 too (destination contributes 0).
__global__ void __launch_bounds__(128) scale_nn_kernel(const uchar4* __restrict__ src, int sw, int sh, int dw, int dh, uchar4* __restrict__ dst) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= dw || y >= dh) return;
    unsigned long long sx = (2ull * (unsigned)x + 1ull) * (unsigned)sw / (2ull * (unsigned)dw);
    unsigned long long sy = (2ull * (unsigned)y + 1ull) * (unsigned)sh / (2ull * (unsigned)dh);
    dst[(size_t)y * dw + x] = src[(size_t)sy * sw + sx];
}

constexpr int kAnsiCell = 41, kAnsiEol = 5;
// One thread per terminal cell: ESC[48;2;RRR;GGG;BBBm ESC[38;2;RRR;GGG;BBBm U+2584 ; thread x == w writes ESC[0m LF.
__global__ void __launch_bounds__(128) ansi_kernel(const uchar4* __restrict__ img, int w, int rows, unsigned char* __restrict__ out) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (x > w || r >= rows) return;
    unsigned char* o = out + (size_t)r * ((size_t)w * kAnsiCell + kAnsiEol) + (size_t)x * kAnsiCell;
    if (x == w) { o[0] = 0x1b; o[1] = '['; o[2] = '0'; o[3] = 'm'; o[4] = '\n'; return; }
    uchar4 px[2] = {img[(size_t)(2 * r) * w + x], img[(size_t)(2 * r + 1) * w + x]};
#pragma unroll
    for (int h = 0; h < 2; h++) {
        o[0] = 0x1b; o[1] = '['; o[2] = h == 0 ? '4' : '3'; o[3] = '8'; o[4] = ';'; o[5] = '2'; o[6] = ';';
        o += 7;
        unsigned v[3] = {px[h].x, px[h].y, px[h].z};
#pragma unroll
        for (int k = 0; k < 3; k++) {
            o[0] = (unsigned char)('0' + v[k] / 100); o[1] = (unsigned char)('0' + (v[k] / 10) % 10); o[2] = (unsigned char)('0' + v[k] % 10);
            o[3] = k < 2 ? ';' : 'm';
            o += 4;
        }
    }
    o[0] = 0xE2; o[1] = 0x96; o[2] = 0x84;
}

// Generator parity probe (single thread).
__global__ void rng_dump_kernel(int kind, unsigned long long idx, unsigned long long seed, double radius, int n, double* out, int indisc_variant, int unitvec_variant) {
    __shared__ ZigTables zig;
    zig_load(&zig, threadIdx.x, blockDim.x);
    __syncthreads();
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    Pcg r = pcg_new_idx(idx, seed);
    for (int i = 0; i < n; i++) {
        if (kind == 0) out[i] = __longlong_as_double((long long)pcg_u64(r));
        else if (kind == 1) out[i] = pcg_f64(r);
        else if (kind == 2) out[i] = pcg_norm(r, &zig);
        else if (kind == 3) { V3<double> v = pcg_unit_vector(r, &zig, unitvec_variant); out[3 * i] = v.x; out[3 * i + 1] = v.y; out[3 * i + 2] = v.z; }
        else { double x, y; pcg_in_disc(r, radius, x, y, indisc_variant); out[2 * i] = x; out[2 * i + 1] = y; }
    }
}

// Arithmetic parity probe: the device's out-of-line IEEE operations on caller-supplied operands (checked on the host against
// its own IEEE division / square root). kind 0: div3_f64(a[i], a[i+1], a[i+2], b[i]) (3 per element, indices mod n);
// 1: div_by_rcp(a[i], b[i], rcp_refined(b[i])); 2: sqrt_f64(a[i]); 3: div_f64(a[i], b[i]).
__global__ void arith_probe_kernel(int kind, const double* a, const double* b, int n, double* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (kind == 0) {
        const V3<double> q = div3_f64(a[i], a[(i + 1) % n], a[(i + 2) % n], b[i]);
        out[3 * i] = q.x; out[3 * i + 1] = q.y; out[3 * i + 2] = q.z;
    } else if (kind == 1) out[i] = div_by_rcp(a[i], b[i], rcp_refined(b[i]));
    else if (kind == 2) out[i] = sqrt_f64(a[i]);
    else out[i] = div_f64(a[i], b[i]);
}

__global__ void srgb_kernel(const double* x, int n, const double* thr, unsigned char* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = linear_to_srgb(thr, x[i]);
}

// Hot-loop-only probe: the sphere-test loop of trace_kernel with nothing around it (no shading, no regeneration,
// rays that certainly miss), to measure the ceiling the loop itself can reach at a given occupancy.
template <typename T, bool FMA, int TPB, int MINB>
__global__ void __launch_bounds__(TPB, MINB) hotloop_probe_kernel(const DevScene<T> S, int iters, double* sink) {
    typedef typename Vec4T<T>::type T4;
    constexpr int CH = TRAY_CH;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T4* sgeo = reinterpret_cast<T4*>(smem_raw);
    for (int i = threadIdx.x; i < S.n_pad; i += TPB) sgeo[i] = S.geo[i];
    __syncthreads();
    const int gid = blockIdx.x * TPB + threadIdx.x;
    // upward rays from high above the scene: h < 0 and c > 0 for every sphere
    T ox = T(0.001) * T(gid & 1023), oy = T(5000), oz = T(0.002) * T(gid >> 10), dx = T(0.01), dy = T(1), dz = T(0.02);
    const T a = dx * dx + dy * dy + dz * dz;
    unsigned acc = 0;
    for (int it = 0; it < iters; it++) {
        unsigned mask_prev = 0;
#pragma unroll 1
        for (int i = 0; i < S.n_pad; i += CH) {
            unsigned mask = 0;
#pragma unroll
            for (int u = 0; u < CH; u++) {
                T4 g = sgeo[i + u];
                T h, c, disc;
                sphere_terms<T, FMA>(ox, oy, oz, dx, dy, dz, a, g.x, g.y, g.z, g.w, h, c, disc);
                mask = __funnelshift_l((unsigned)miss_bits(h, c, disc), mask, 1);
            }
            mask = ~mask & ((1u << CH) - 1u);
            if (mask_prev) acc += mask_prev;
            mask_prev = mask;
        }
        acc += mask_prev;
        ox += T(1e-9);  // keep the compiler from hoisting the loop
    }
    if (acc == 0xdeadbeefu) sink[0] = (double)acc;
}

// Issue-bound pipe peaks (roofline denominators): 8 independent chains per thread.
template <int KIND>
__global__ void __launch_bounds__(256) peak_kernel(double* sink, int iters, double seedv) {
    if (KIND == 5) {  // packed FFMA2 (two fp32 FMAs per instruction, sm_100+)
        float2 acc[8];
        float2 m = make_float2(1.0000001f, 0.9999999f), b = make_float2((float)seedv, (float)seedv * 2);
#pragma unroll
        for (int k = 0; k < 8; k++) acc[k] = make_float2((float)(threadIdx.x + k), (float)k);
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int k = 0; k < 8; k++) acc[k] = __ffma2_rn(acc[k], m, b);
        }
        float s = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) s += acc[k].x + acc[k].y;
        if (s == 12345.678f) sink[0] = s;
    } else if (KIND == 2) {
        float acc[8];
        float m = 1.0000001f, b = (float)seedv;
#pragma unroll
        for (int k = 0; k < 8; k++) acc[k] = (float)(threadIdx.x + k);
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int k = 0; k < 8; k++) acc[k] = fmaf(acc[k], m, b);
        }
        float s = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) s += acc[k];
        if (s == 12345.678f) sink[0] = s;
    } else {
        double acc[8];
        double m = 1.0000000001, b = seedv;
#pragma unroll
        for (int k = 0; k < 8; k++) acc[k] = (double)(threadIdx.x + k);
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (KIND == 0) acc[k] = fma(acc[k], m, b);
                else acc[k] = acc[k] * m + b;  // -fmad=false: one DMUL + one DADD
            }
        }
        double s = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) s += acc[k];
        if (s == 12345.678) sink[0] = s;
    }
}

}  // namespace tray
