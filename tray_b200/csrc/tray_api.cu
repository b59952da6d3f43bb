// tray_api.cu -- host side of libtraycuda.so: the C ABI declared in include/tray_cuda.h.
// Owns device memory, streams and events; copies every caller buffer before returning.
// There is deliberately NO CPU fallback: without a usable sm_100 device every entry point fails.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/tray_cuda.h"
#include "tray_kernels.cuh"
#include "tray_png.cuh"
#include "tray_wavefront.cuh"
#include "tray_lbvh.cuh"

using namespace tray;

namespace {

thread_local std::string g_init_error;

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            char buf_[512];                                                                        \
            snprintf(buf_, sizeof buf_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            throw std::runtime_error(buf_);                                                        \
        }                                                                                          \
    } while (0)

// ---- host copies of the deterministic log/exp used to build the sRGB threshold table --------
// (same FreeBSD-msun forms as tray_device.cuh; host build uses -ffp-contract=off)
double h_log(double x) {
    const double Ln2Hi = 6.93147180369123816490e-01, Ln2Lo = 1.90821492927058770002e-10;
    const double L1 = 6.666666666666735130e-01, L2 = 3.999999999940941908e-01, L3 = 2.857142874366239149e-01,
                 L4 = 2.222219843214978396e-01, L5 = 1.818357216161805012e-01, L6 = 1.531383769920937332e-01,
                 L7 = 1.479819860511658591e-01;
    int ki;
    double f1 = std::frexp(x, &ki);
    if (f1 < 0.70710678118654752440) { f1 *= 2; ki--; }
    double f = f1 - 1, k = (double)ki;
    double s = f / (2 + f), s2 = s * s, s4 = s2 * s2;
    double t1 = s2 * (L1 + s4 * (L3 + s4 * (L5 + s4 * L7)));
    double t2 = s4 * (L2 + s4 * (L4 + s4 * L6));
    double R = t1 + t2, hfsq = 0.5 * f * f;
    return k * Ln2Hi - ((hfsq - (s * (hfsq + R) + k * Ln2Lo)) - f);
}
double h_exp(double x) {
    const double Ln2Hi = 6.93147180369123816490e-01, Ln2Lo = 1.90821492927058770002e-10, Log2e = 1.44269504088896338700e+00;
    const double NearZero = 1.0 / (1 << 28);
    const double P1 = 1.66666666666666657415e-01, P2 = -2.77777777770155933842e-03, P3 = 6.61375632143793436117e-05,
                 P4 = -1.65339022054652515390e-06, P5 = 4.13813679705723846039e-08;
    if (-NearZero < x && x < NearZero) return 1 + x;
    int k = 0;
    if (x < 0) k = (int)(Log2e * x - 0.5);
    else if (x > 0) k = (int)(Log2e * x + 0.5);
    double hi = x - (double)k * Ln2Hi, lo = (double)k * Ln2Lo;
    double r = hi - lo, t = r * r;
    double c = r - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
    double y = 1 - ((lo - (r * c) / (2 - c)) - hi);
    return std::ldexp(y, k);
}
// tcolor.LinearToSrgb (third-party, call site ray/vec3.go:175-177; pinned by ray/vec3_test.go:264-289)
int h_linear_to_srgb(double x) {
    if (!(x > 0)) return 0;
    if (x >= 1) return 255;
    double s = x <= 0.0031308 ? 12.92 * x : 1.055 * h_exp((1 / 2.4) * h_log(x)) - 0.055;
    return (int)std::round(255 * s);
}
// thr[k] = smallest double x with LinearToSrgb(x) >= k (bisection on the bit pattern; the map is monotone).
void build_srgb_thresholds(double* thr) {
    thr[0] = 0.0;
    for (int k = 1; k <= 255; k++) {
        uint64_t lo = 0, hi;  // lo: conv < k ; hi: conv >= k
        double one = 1.0;
        memcpy(&hi, &one, 8);
        while (hi - lo > 1) {
            uint64_t mid = lo + (hi - lo) / 2;
            double x;
            memcpy(&x, &mid, 8);
            if (h_linear_to_srgb(x) >= k) hi = mid; else lo = mid;
        }
        memcpy(&thr[k], &hi, 8);
    }
}

struct Device {
    int dev = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    std::vector<cudaEvent_t> ev_pool;  // (start, stop) pairs around every trace launch of the last render
    int ev_used = 0;
    // scene
    int n = 0, n_pad = 0;
    double4* geo_d = nullptr; float4* geo_f = nullptr;
    double* radius_d = nullptr; float* radius_f = nullptr;
    uint8_t* kind = nullptr; double4* params = nullptr;
    float4* fpair = nullptr; float filt_mc = 0, filt_r2max = 0;
    // two-level cluster tables (cluster_scan)
    float4* cl_blob = nullptr; size_t cap_cl_blob = 0;
    int cl_blob_f4 = 0, cl_off_box2 = 0, cl_off_box1 = 0, cl_off_box0 = 0, cl_off_ids = 0, cl_real_groups = 0, cl_always_groups = 0;
    unsigned cl_always_last = 0; float cl_r = 0; bool cl_present = false;
    BvhNode* bvh = nullptr; int* bvh_leaf_ids = nullptr; int* bvh_always = nullptr; int bvh_n_always = 0; double bvh_extent = 0;
    bool bvh_present = false;
    // the scene buffers are kept between uploads and only grown (cudaFree / cudaMalloc per keypress cost up to tens of ms)
    size_t cap_geo_d = 0, cap_geo_f = 0, cap_radius_d = 0, cap_radius_f = 0, cap_kind = 0, cap_params = 0, cap_fpair = 0,
           cap_bvh = 0, cap_leaf = 0, cap_always = 0;
    // work buffers (grown on demand)
    double* scratch = nullptr; size_t scratch_cap = 0;  // doubles
    uint8_t* rgba = nullptr; size_t rgba_cap = 0;       // bytes
    double* hdr = nullptr; size_t hdr_cap = 0;          // doubles
    unsigned long long* counters = nullptr; int counters_cap = 0;
    unsigned short* stk_g = nullptr; size_t stk_cap = 0;  // regroup layout: attenuation stacks [level][slot]
    CamRay* gen = nullptr; size_t gen_cap = 0;             // camera rays of a pass (camera_ray_kernel)
    // LBVH build scratch (tray_lbvh.cuh), kept between uploads
    unsigned long long* lbvh_keys = nullptr; size_t lbvh_keys_cap = 0;
    int* lbvh_int = nullptr; size_t lbvh_int_cap = 0;
    int2* lbvh_int2 = nullptr; size_t lbvh_int2_cap = 0;
    // wavefront layout: path records (two arrays), per-material queues, free stack slots, counters
    WfRec* wf_rec = nullptr; size_t wf_rec_cap = 0;
    unsigned* wf_queue = nullptr; size_t wf_queue_cap = 0;
    unsigned* wf_free = nullptr; size_t wf_free_cap = 0;
    WfCounters* wf_cnt = nullptr; size_t wf_cnt_cap = 0;
    unsigned* wf_host = nullptr;  // pinned: n_in read back every few bounces
    // tray_present work buffers, kept between calls
    double4* pr_tmp = nullptr; size_t pr_tmp_cap = 0;
    uchar4* pr_small = nullptr; size_t pr_small_cap = 0;
    unsigned char* pr_ansi = nullptr; size_t pr_ansi_cap = 0;
    // PNG encoder work buffers (tray_encode_png), kept between calls
    unsigned char* png_filt = nullptr; size_t png_filt_cap = 0;
    unsigned char* png_out = nullptr; size_t png_out_cap = 0;
    unsigned* png_hist = nullptr; size_t png_hist_cap = 0;
    unsigned* png_piece = nullptr; size_t png_piece_cap = 0;
    unsigned long long* png_adler = nullptr; size_t png_adler_cap = 0;
    PngBlock* png_blocks = nullptr; size_t png_blocks_cap = 0;
    PngTotals* png_tot = nullptr; size_t png_tot_cap = 0;
    unsigned long long* stats = nullptr;     // [0] segments [1] depth exhausted [2] progress samples
    double* srgb_thr = nullptr;
    uint8_t* pinned = nullptr; size_t pinned_cap = 0;
    // tray_progress: a side stream and a pinned word per device (another thread peeks at the counters of a render in flight)
    cudaStream_t peek_stream = nullptr;
    unsigned long long* peek_host = nullptr;
    // in-context sample split: can device 0 read this device's memory (peer access over NVLink)? If not, its partial sums
    // are staged into a buffer on device 0 with cudaMemcpyPeerAsync before the combine kernel.
    bool peer_of_root = true;
    double* stage = nullptr; size_t stage_cap = 0;   // lives on device 0, one per non-peer device
    // last render bookkeeping
    std::vector<int> local_rows;  // global row of each local row
    int passes = 0;
    bool timed = false;           // ev_begin / ev_end were recorded by the last render
};

// Frees a temporary device allocation on every exit path (the parity probes allocate per call).
struct DevTmp {
    void* p = nullptr;
    explicit DevTmp(size_t bytes) { CK(cudaMalloc(&p, bytes ? bytes : 1)); }
    ~DevTmp() { cudaFree(p); }
    DevTmp(const DevTmp&) = delete;
    DevTmp& operator=(const DevTmp&) = delete;
    template <typename P> P* as() const { return static_cast<P*>(p); }
};

cudaEvent_t next_event(Device& d) {
    if (d.ev_used == (int)d.ev_pool.size()) {
        cudaEvent_t e;
        CK(cudaEventCreate(&e));
        d.ev_pool.push_back(e);
    }
    return d.ev_pool[d.ev_used++];
}

template <typename P>
void grow(P*& ptr, size_t& cap, size_t need) {
    if (need <= cap) return;
    if (ptr) CK(cudaFree(ptr));
    ptr = nullptr; cap = 0;
    CK(cudaMalloc(&ptr, need * sizeof(P)));
    cap = need;
}

}  // namespace

struct tray_ctx {
    std::vector<Device> devs;
    std::string err;
    std::mutex mu;
    // The BVH of a scene that the cluster walk serves (TRAY_ACCEL_AUTO never needs it) is built on first use, from these copies
    // of the uploaded centres and radii: tray_scene_upload is on the per-keypress path (0.04 ms of 0.18 at 485 spheres, 2 ms at 10 001).
    std::vector<double> bvh_cx, bvh_cy, bvh_cz, bvh_r;
    bool bvh_pending = false;
    bool have_scene = false;
    std::vector<double4> host_geo_d;  // host copy of the padded table (kernel-parameter path)
    std::vector<float4> host_geo_f;
    double bg_a[3] = {1, 1, 1}, bg_b[3] = {0.4, 0.65, 1.0};
    // last render
    int width = 0, height = 0, y0 = 0, y1 = 0;
    bool have_image = false, have_hdr = false;
    int split_mode = 0;
    int bvh_build = TRAY_BVH_BUILD_AUTO;  // tray_configure(TRAY_CFG_BVH_BUILD)
    int bvh_built_on_device = 0;          // what the last upload did (tray_query)
    int cluster_unfilterable = 0;         // spheres of the scene the fp32 filter cannot bound (every ray tests them exactly)
    int indisc_variant = 0, unitvec_variant = 0;  // tray_configure(TRAY_CFG_INDISC / TRAY_CFG_UNITVEC)
    int debug_fault = 0;                          // bounds-check builds only (negative control)
    bool hdr_is_sums = false;          // last render left raw colour sums (sums_mode != 0)
    uint64_t sums_samples = 0;         // samples per pixel accumulated in them
    int sums_key[6] = {0, 0, 0, 0, 0, 0};  // geometry the sums belong to (w, h, y0, y1, shard_index, shard_count)
    std::atomic<uint64_t> progress_base{0};
    std::atomic<int> progress_spp{1};
    std::atomic<bool> rendering{false};
    std::mutex peek_mu;                // tray_progress callers take turns on the per-device pinned words
};

namespace {

int fail(tray_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg; else g_init_error = msg;
    return code;
}

void free_scene(Device& d) {
    cudaSetDevice(d.dev);
    cudaFree(d.geo_d); cudaFree(d.geo_f); cudaFree(d.radius_d); cudaFree(d.radius_f); cudaFree(d.kind); cudaFree(d.params);
    cudaFree(d.fpair); d.fpair = nullptr;
    cudaFree(d.cl_blob); d.cl_blob = nullptr; d.cap_cl_blob = 0; d.cl_present = false;
    cudaFree(d.bvh); cudaFree(d.bvh_leaf_ids); cudaFree(d.bvh_always);
    d.bvh = nullptr; d.bvh_leaf_ids = nullptr; d.bvh_always = nullptr; d.bvh_n_always = 0; d.bvh_present = false;
    d.cap_geo_d = d.cap_geo_f = d.cap_radius_d = d.cap_radius_f = d.cap_kind = d.cap_params = d.cap_fpair = d.cap_bvh = d.cap_leaf = d.cap_always = 0;
    d.geo_d = nullptr; d.geo_f = nullptr; d.radius_d = nullptr; d.radius_f = nullptr; d.kind = nullptr; d.params = nullptr;
}

template <typename T>
void fill_cluster(DevScene<T>& s, const Device& d) {
    s.cl_blob = d.cl_present ? d.cl_blob : nullptr; s.cl_blob_f4 = d.cl_blob_f4;
    s.cl_off_box2 = d.cl_off_box2; s.cl_off_box1 = d.cl_off_box1; s.cl_off_box0 = d.cl_off_box0; s.cl_off_ids = d.cl_off_ids;
    s.cl_real_groups = d.cl_real_groups; s.cl_always_groups = d.cl_always_groups; s.cl_always_last = d.cl_always_last; s.cl_r = d.cl_r;
}
template <typename T> DevScene<T> dev_scene(const tray_ctx* ctx, const Device& d);
template <> DevScene<double> dev_scene<double>(const tray_ctx* ctx, const Device& d) {
    DevScene<double> s;
    s.n = d.n; s.n_pad = d.n_pad; s.geo = d.geo_d; s.radius = d.radius_d; s.kind = d.kind; s.params = d.params;
    s.fpair = d.fpair; s.filt_mc = d.filt_mc; s.filt_r2max = d.filt_r2max;
    fill_cluster(s, d);
#ifdef TRAY_BOUNDS_CHECK
    if (ctx->debug_fault) s.cl_blob_f4 = std::max(0, s.cl_blob_f4 - 256);  // the kernel then stages (and checks against) too small a table
#endif
    s.unitvec_variant = ctx->unitvec_variant;
    s.bvh = d.bvh_present ? d.bvh : nullptr; s.bvh_leaf_ids = d.bvh_leaf_ids; s.bvh_always = d.bvh_always; s.bvh_n_always = d.bvh_n_always; s.bvh_extent = d.bvh_extent;
    for (int i = 0; i < 3; i++) { s.bg_a[i] = ctx->bg_a[i]; s.bg_b[i] = ctx->bg_b[i]; }
    return s;
}
template <> DevScene<float> dev_scene<float>(const tray_ctx* ctx, const Device& d) {
    DevScene<float> s;
    s.n = d.n; s.n_pad = d.n_pad; s.geo = d.geo_f; s.radius = d.radius_f; s.kind = d.kind; s.params = d.params;
    s.fpair = d.fpair; s.filt_mc = d.filt_mc; s.filt_r2max = d.filt_r2max;
    fill_cluster(s, d);
    s.unitvec_variant = ctx->unitvec_variant;
    s.bvh = nullptr; s.bvh_leaf_ids = nullptr; s.bvh_always = nullptr; s.bvh_n_always = 0; s.bvh_extent = 0;
    for (int i = 0; i < 3; i++) { s.bg_a[i] = ctx->bg_a[i]; s.bg_b[i] = ctx->bg_b[i]; }
    return s;
}

DevCamera dev_camera(const tray_camera* c, int indisc_variant = 0) {
    DevCamera d;
    d.indisc_variant = indisc_variant;
    for (int i = 0; i < 3; i++) {
        d.pos[i] = c->position[i]; d.p00[i] = c->pixel00[i]; d.px[i] = c->pixel_x[i]; d.py[i] = c->pixel_y[i];
        d.du[i] = c->defocus_u[i]; d.dv[i] = c->defocus_v[i];
    }
    d.aperture = c->aperture; d.focus_distance = c->focus_distance; d.focal_length = c->focal_length;
    d.focus_time = c->focus_distance / c->focal_length;  // (host build: -ffp-contract=off, IEEE double division like the device's)
    return d;
}

#ifndef TRAY_TPB
#define TRAY_TPB 128
#endif
#ifndef TRAY_MINB
#define TRAY_MINB 5
#endif
constexpr int kTPB = TRAY_TPB;        // threads per CTA of the trace kernel
constexpr int kMinBlocks = TRAY_MINB; // resident CTAs per SM the register allocation targets
constexpr size_t kSmemBudget = 200 * 1024;
constexpr int kSmemClusterMax = 2048;  // up to here the cluster tables are staged into shared memory (two levels); above: cluster_scan_big

// Regroup layout (pre-filter kernel only): same launch shape, plus the exchange area in shared memory and the
// attenuation stacks in global memory (max_depth x lanes x 2 bytes, L2 resident).
template <typename T, bool FMA, int GEO>
void launch_trace_regroup(Device& d, TraceArgs A, const DevScene<T>& S, size_t smem) {
#ifndef TRAY_FILTER_MINB
#define TRAY_FILTER_MINB 4
#endif
    auto k = trace_kernel<T, FMA, kTPB, TRAY_FILTER_MINB, GEO, true>;
    smem = ((smem + 15) & ~(size_t)15) + sizeof(RegroupBuf<kTPB>);  // (smem already holds the 16 bytes of alignment slack)
    int bps = 0;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k, kTPB, smem));
    if (bps < 1) bps = 1;
    const unsigned grid = (unsigned)(d.num_sms * bps);
    A.n_slots = grid * kTPB;
    grow(d.stk_g, d.stk_cap, (size_t)A.n_slots * (size_t)A.max_depth);
    A.stk_g = d.stk_g;
    GeoArg<T, GEO> none{};
    k<<<grid, kTPB, smem, d.stream>>>(A, S, none);
    CK(cudaGetLastError());
}

template <typename T, bool FMA, int GEO>
void launch_trace_geo(const Device& d, const TraceArgs& A, const DevScene<T>& S, const GeoArg<T, GEO>& GP, size_t smem) {
    // register budget: the pre-filter kernels keep both the fp32 filter state and the fp64 ray live: 128 regs x 16 warps/SM
    // measured best (140.8 vs 143.3 ms); the pure-fp64 kernels prefer 96 regs x 20 warps/SM (211.8 vs 216.7 ms).
#ifndef TRAY_FILTER_MINB
#define TRAY_FILTER_MINB 4
#endif
#ifndef TRAY_CLUSTER_MINB
#define TRAY_CLUSTER_MINB 4
#endif
#ifndef TRAY_FP32_MINB
#define TRAY_FP32_MINB 6
#endif
    constexpr int minb = GEO == kGeoFilter ? TRAY_FILTER_MINB
                       : ((GEO == kGeoCluster || GEO == kGeoClusterBig) ? (sizeof(T) == 4 ? TRAY_FP32_MINB : TRAY_CLUSTER_MINB) : kMinBlocks);
    auto k = trace_kernel<T, FMA, kTPB, minb, GEO>;
    if constexpr (GEO == kGeoCluster && ((sizeof(T) == 8 && !FMA) || sizeof(T) == 4)) {  // (the strict default and the fp32 fast path)
        // small passes (config 1: 380 samples per resident warp) take the counter in steps of 32 instead of 128 samples: a separate
        // instantiation, so that the kernel of the large frames stays byte for byte what it was (0.59 -> 0.49 ms per config-1 frame)
        if (A.n_samples / ((unsigned long long)d.num_sms * 16ull) < 8ull * kBatch) k = trace_kernel<T, FMA, kTPB, minb, GEO, false, kBatchSmall>;
    }
    int bps = 0;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k, kTPB, smem));
    if (bps < 1) bps = 1;
    k<<<d.num_sms * bps, kTPB, smem, d.stream>>>(A, S, GP);
    CK(cudaGetLastError());
}

// which closest-hit structure a render uses (results are identical for all of them)
enum { kUseLinear = 0, kUseBvh = 1, kUseCluster = 2 };

template <typename T, bool FMA>
struct TraceLaunch {
    static void run(Device& d, const TraceArgs& A, const DevScene<T>& S, const void* host_geo, bool filter = false, int use = kUseLinear, bool regroup = false) {
        typedef typename Vec4T<T>::type T4;
        const size_t tail = sizeof(ZigTables) + (size_t)kCand * kTPB * sizeof(uint16_t) + 16 + kGenPoolBytes<kTPB>;
        const size_t geo_bytes = (size_t)S.n_pad * sizeof(T4);
        if constexpr (sizeof(T) == 8) {
            if (use == kUseBvh) {
                GeoArg<T, kGeoBVH> none{};
                launch_trace_geo<T, FMA, kGeoBVH>(d, A, S, none, tail);
                return;
            }
        }
        if (filter && use == kUseCluster && S.cl_blob && S.n <= kSmemClusterMax && (size_t)S.cl_blob_f4 * 16 + tail <= kSmemBudget) {
            if (regroup) { launch_trace_regroup<T, FMA, kGeoCluster>(d, A, S, (size_t)S.cl_blob_f4 * 16 + tail); return; }
            GeoArg<T, kGeoCluster> none{};
            launch_trace_geo<T, FMA, kGeoCluster>(d, A, S, none, (size_t)S.cl_blob_f4 * 16 + tail);
            return;
        }
        if (filter && use == kUseCluster && S.cl_blob) {  // large scene: three levels, tables in global memory (plain layout only)
            GeoArg<T, kGeoClusterBig> none{};
            launch_trace_geo<T, FMA, kGeoClusterBig>(d, A, S, none, tail);
            return;
        }
        if (filter && S.fpair && (size_t)S.n_pad * 16 + tail <= kSmemBudget) {
            if (regroup) { launch_trace_regroup<T, FMA, kGeoFilter>(d, A, S, (size_t)S.n_pad * 16 + tail); return; }
            GeoArg<T, kGeoFilter> none{};
            launch_trace_geo<T, FMA, kGeoFilter>(d, A, S, none, (size_t)S.n_pad * 16 + tail);
            return;
        }
        if (TRAY_PARAM_GEO && S.n_pad <= kParamSpheres) {
            // small scene: the table travels in the kernel parameters (constant bank, uniform loads)
            static thread_local GeoArg<T, kGeoParam> gp;
            memcpy(gp.g, host_geo, geo_bytes);
            for (int i = S.n_pad; i < kParamSpheres; i++) gp.g[i] = ((const T4*)host_geo)[S.n_pad - 1];
            launch_trace_geo<T, FMA, kGeoParam>(d, A, S, gp, tail);
        } else if (geo_bytes + tail <= kSmemBudget) {
            GeoArg<T, kGeoShared> none{};
            launch_trace_geo<T, FMA, kGeoShared>(d, A, S, none, geo_bytes + tail);
        } else {
            GeoArg<T, kGeoGlobal> none{};
            launch_trace_geo<T, FMA, kGeoGlobal>(d, A, S, none, tail);
        }
    }
};

// Wavefront layout: the host drives one bounce at a time (intersect, shade, regenerate, advance) until no path is left.
constexpr unsigned kWfCap = 1u << 21;  // paths in flight (records: 2 x 201 MB)
int launch_trace_wavefront(Device& d, const TraceArgs& A, const DevScene<double>& S) {
    const unsigned cap = (unsigned)std::min<unsigned long long>(kWfCap, (A.n_samples + kWfTPB - 1) / kWfTPB * kWfTPB);
    grow(d.wf_rec, d.wf_rec_cap, (size_t)2 * cap);
    grow(d.wf_queue, d.wf_queue_cap, (size_t)4 * cap);
    grow(d.wf_free, d.wf_free_cap, (size_t)cap);
    grow(d.wf_cnt, d.wf_cnt_cap, (size_t)1);
    grow(d.stk_g, d.stk_cap, (size_t)cap * (size_t)A.max_depth);
    if (!d.wf_host) CK(cudaHostAlloc(&d.wf_host, sizeof(unsigned), cudaHostAllocDefault));
    WfArgs W;
    W.A = A; W.A.stk_g = d.stk_g; W.A.n_slots = cap;
    W.cur = d.wf_rec; W.next = d.wf_rec + cap; W.queue = d.wf_queue; W.freelist = d.wf_free; W.cnt = d.wf_cnt; W.cap = cap;
    const unsigned grid = cap / kWfTPB;
    const size_t smem = (size_t)S.n_pad * 16 + 256 + (size_t)kCand * kWfTPB * sizeof(uint16_t);
    CK(cudaFuncSetAttribute(wf_intersect_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int bps = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, wf_intersect_kernel<false>, kWfTPB, smem));
    const unsigned grid_int = (unsigned)std::min<unsigned long long>((unsigned long long)d.num_sms * std::max(1, bps), (cap + kWfTPB - 1) / kWfTPB);
    int launches = 0;
    wf_init_kernel<<<(cap + 255) / 256, 256, 0, d.stream>>>(W);
    wf_regen_kernel<<<grid, kWfTPB, 0, d.stream>>>(W);
    wf_advance_kernel<<<1, 1, 0, d.stream>>>(W.cnt);
    std::swap(W.cur, W.next);
    launches += 3;
    for (int bounce = 1;; bounce++) {
        wf_intersect_kernel<false><<<grid_int, kWfTPB, smem, d.stream>>>(W, S);
        wf_shade_kernel<<<grid, kWfTPB, 0, d.stream>>>(W, S);
        wf_regen_kernel<<<grid, kWfTPB, 0, d.stream>>>(W);
        wf_advance_kernel<<<1, 1, 0, d.stream>>>(W.cnt);
        std::swap(W.cur, W.next);
        launches += 4;
        if (bounce % 8 == 0) {
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(d.wf_host, &W.cnt->n_in, sizeof(unsigned), cudaMemcpyDeviceToHost, d.stream));
            CK(cudaStreamSynchronize(d.stream));
            if (*d.wf_host == 0) break;
        }
        if (bounce > 100000000) throw std::runtime_error("wavefront: no progress");
    }
    CK(cudaGetLastError());
    return launches;
}

#ifndef TRAY_DEFAULT_LAYOUT
#define TRAY_DEFAULT_LAYOUT TRAY_LAYOUT_PLAIN  // config 2, with the camera rays generated ahead: plain 99.7 ms, regroup 101.8, wavefront 120.9
#endif
// AUTO: the cluster walk while the scene is small enough for its tables (32768 spheres) and the filter can bound nearly all of
// it, the per-lane BVH otherwise; BRUTE: the linear scan in the reference's Scene.Hit order.
constexpr int kAutoClusterMax = 32768, kAutoClusterUnfilterable = 64;
int closest_hit_structure(const tray_ctx* ctx, const Device& d, int accel) {
    const bool cluster_ok = d.cl_present;
    if (accel == TRAY_ACCEL_BVH) return kUseBvh;
    if (accel == TRAY_ACCEL_CLUSTER) return cluster_ok ? kUseCluster : kUseLinear;
    if (accel == TRAY_ACCEL_BRUTE) return kUseLinear;
    if (cluster_ok && d.n <= kAutoClusterMax && ctx->cluster_unfilterable <= kAutoClusterUnfilterable) return kUseCluster;
    return d.n > kSmemClusterMax ? kUseBvh : kUseLinear;
}

// The wavefront layout exists for the strict fp64 linear scan only; everything else runs the megakernel.
bool wavefront_runs(const tray_ctx* ctx, const Device& d, int precision, int accel, int layout) {
    const bool bvh = closest_hit_structure(ctx, d, accel) == kUseBvh;
    const size_t tail = sizeof(ZigTables) + (size_t)kCand * kTPB * sizeof(uint16_t);
    return layout == TRAY_LAYOUT_WAVEFRONT && precision == TRAY_FP64_STRICT && !bvh && (size_t)d.n_pad * 16 + tail <= kSmemBudget;
}

int launch_trace(const tray_ctx* ctx, Device& d, const TraceArgs& A, int precision, int accel, int layout) {
    const int use = closest_hit_structure(ctx, d, accel);
    const bool auto_layout = layout == TRAY_LAYOUT_AUTO;
    if (auto_layout) layout = TRAY_DEFAULT_LAYOUT;
    if (wavefront_runs(ctx, d, precision, accel, layout)) return launch_trace_wavefront(d, A, dev_scene<double>(ctx, d));
    if (layout == TRAY_LAYOUT_WAVEFRONT) layout = TRAY_LAYOUT_REGROUP;
    const bool regroup = layout == TRAY_LAYOUT_REGROUP;
    // the all-fp64 kernels (FMA, STRICT_BRUTE) scan linearly or walk the BVH; the pre-filter kernels add the cluster structure
    const int use64 = use == kUseBvh ? kUseBvh : kUseLinear;
    if (precision == TRAY_FP64_FMA) TraceLaunch<double, true>::run(d, A, dev_scene<double>(ctx, d), ctx->host_geo_d.data(), false, use64);
    else if (precision == TRAY_FP64_STRICT) TraceLaunch<double, false>::run(d, A, dev_scene<double>(ctx, d), ctx->host_geo_d.data(), true, use, regroup);
    else if (precision == TRAY_FP64_STRICT_BRUTE) TraceLaunch<double, false>::run(d, A, dev_scene<double>(ctx, d), ctx->host_geo_d.data(), false, use64);
    // fp32 fast path: the same conservative pre-filter (it proves a miss in exact arithmetic), survivors tested in fp32
    else TraceLaunch<float, true>::run(d, A, dev_scene<float>(ctx, d), ctx->host_geo_f.data(), true, use == kUseBvh ? kUseLinear : use, regroup);
    return 1;
}

// ---- host BVH build: median split on the longest centroid axis, <= 4 spheres per leaf ----------------------
struct HostBvh {
    std::vector<BvhNode> nodes;
    std::vector<int> leaf_ids, always;
    double extent = 0;
    int max_depth = 0;
};

void bvh_bounds(const tray_scene_desc* sc, const int* ids, int n, double lo[3], double hi[3]) {
    for (int k = 0; k < 3; k++) { lo[k] = std::numeric_limits<double>::infinity(); hi[k] = -lo[k]; }
    for (int j = 0; j < n; j++) {
        int i = ids[j];
        const double c[3] = {sc->cx[i], sc->cy[i], sc->cz[i]};
        const double r = std::fabs(sc->radius[i]);
        for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], c[k] - r); hi[k] = std::max(hi[k], c[k] + r); }
    }
    for (int k = 0; k < 3; k++) {  // pad outwards: 2^-30 relative (+ tiny absolute), far above any fp64 rounding in the tests
        double m = std::max(std::fabs(lo[k]), std::fabs(hi[k])) + 1.0;
        lo[k] -= m * 9.4e-10; hi[k] += m * 9.4e-10;
    }
}

int bvh_build_rec(const tray_scene_desc* sc, HostBvh& b, std::vector<int>& ids, int first, int n, int depth) {
    int me = (int)b.nodes.size();
    b.nodes.push_back(BvhNode{});
    b.max_depth = std::max(b.max_depth, depth);
    double lo[3], hi[3];
    bvh_bounds(sc, ids.data() + first, n, lo, hi);
    for (int k = 0; k < 3; k++) { b.nodes[me].lo[k] = lo[k]; b.nodes[me].hi[k] = hi[k]; b.extent = std::max(b.extent, std::max(std::fabs(lo[k]), std::fabs(hi[k]))); }
    if (n <= 4) {
        std::sort(ids.begin() + first, ids.begin() + first + n);
        b.nodes[me].left = -((int)b.leaf_ids.size() + 1);
        b.nodes[me].right = n;
        b.nodes[me].axis = 0;
        for (int j = 0; j < n; j++) b.leaf_ids.push_back(ids[first + j]);
        return me;
    }
    double clo[3] = {1e300, 1e300, 1e300}, chi[3] = {-1e300, -1e300, -1e300};
    for (int j = 0; j < n; j++) {
        int i = ids[first + j];
        const double c[3] = {sc->cx[i], sc->cy[i], sc->cz[i]};
        for (int k = 0; k < 3; k++) { clo[k] = std::min(clo[k], c[k]); chi[k] = std::max(chi[k], c[k]); }
    }
    int axis = 0;
    if (chi[1] - clo[1] > chi[axis] - clo[axis]) axis = 1;
    if (chi[2] - clo[2] > chi[axis] - clo[axis]) axis = 2;
    const double* key = axis == 0 ? sc->cx : (axis == 1 ? sc->cy : sc->cz);
    int half = n / 2;
    std::nth_element(ids.begin() + first, ids.begin() + first + half, ids.begin() + first + n,
                     [key](int x, int y) { return key[x] < key[y] || (key[x] == key[y] && x < y); });
    int l = bvh_build_rec(sc, b, ids, first, half, depth + 1);
    int r = bvh_build_rec(sc, b, ids, first + half, n - half, depth + 1);
    b.nodes[me].left = l; b.nodes[me].right = r; b.nodes[me].axis = axis;
    return me;
}

// Which spheres go into the tree (ids) and which are always tested (b.always).
HostBvh bvh_classify(const tray_scene_desc* sc, std::vector<int>& ids);
void bvh_build_host(const tray_scene_desc* sc, HostBvh& b, std::vector<int>& ids) {
    if (!ids.empty()) bvh_build_rec(sc, b, ids, 0, (int)ids.size(), 1);
    if (b.max_depth > 40) throw std::runtime_error("tray_scene_upload: BVH deeper than the traversal stack");
}
HostBvh bvh_classify(const tray_scene_desc* sc, std::vector<int>& ids) {
    HostBvh b;
    int n = sc->n;
    if (n == 0) return b;
    // spheres much larger than the typical one (the r=1000 ground) or with non-finite data stay out of the tree
    std::vector<double> rs(n);
    for (int i = 0; i < n; i++) rs[i] = std::fabs(sc->radius[i]);
    std::vector<double> sorted = rs;
    std::nth_element(sorted.begin(), sorted.begin() + n / 2, sorted.end());
    const double med = sorted[n / 2];
    for (int i = 0; i < n; i++) {
        bool finite = std::isfinite(sc->cx[i]) && std::isfinite(sc->cy[i]) && std::isfinite(sc->cz[i]) && std::isfinite(rs[i]);
        if (!finite || rs[i] > 16.0 * med + 1e-300) b.always.push_back(i); else ids.push_back(i);
    }
    for (int i : ids) {  // max |coordinate| of the tree's boxes (the traversal stops culling for origins absurdly far away)
        double r = rs[i];
        b.extent = std::max(b.extent, std::max(std::fabs(sc->cx[i]) + r, std::max(std::fabs(sc->cy[i]) + r, std::fabs(sc->cz[i]) + r)) * 1.000001 + 1e-9);
    }
    return b;
}

// LBVH on the device (tray_lbvh.cuh). Needs d.geo_d / d.radius_d uploaded. Returns false when the radix tree came out
// deeper than the traversal stack allows (heavily clustered centres): the caller then uses the host's median-split build.
bool bvh_build_device(Device& d, const tray_scene_desc* sc, const std::vector<int>& ids) {
    const int m = (int)ids.size();
    LbvhArgs L;
    L.m = m; L.geo = d.geo_d; L.radius = d.radius_d;
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int i : ids) {
        const double c[3] = {sc->cx[i], sc->cy[i], sc->cz[i]};
        for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], c[k]); hi[k] = std::max(hi[k], c[k]); }
    }
    for (int k = 0; k < 3; k++) { L.lo[k] = lo[k]; L.inv_extent[k] = hi[k] > lo[k] ? 1.0 / (hi[k] - lo[k]) : 0.0; }
    int n_keys = 1;
    while (n_keys < m) n_keys *= 2;
    L.n_keys = n_keys;
    // ids[m] | parent[2m-1] | visits[m] | depth[2m-1]
    grow(d.lbvh_keys, d.lbvh_keys_cap, (size_t)n_keys);
    grow(d.lbvh_int, d.lbvh_int_cap, (size_t)6 * m);
    grow(d.lbvh_int2, d.lbvh_int2_cap, (size_t)2 * m);
    int* d_int = d.lbvh_int;
    CK(cudaMemsetAsync(d_int, 0, sizeof(int) * (size_t)(6 * m), d.stream));
    CK(cudaMemcpyAsync(d_int, ids.data(), sizeof(int) * m, cudaMemcpyHostToDevice, d.stream));
    L.keys = d.lbvh_keys;
    L.ids = d_int; L.parent = d_int + m; L.visits = d_int + 3 * m; L.depth = d_int + 4 * m;
    L.range = d.lbvh_int2; L.child = d.lbvh_int2 + m;
    grow(d.bvh, d.cap_bvh, (size_t)(2 * m - 1));
    grow(d.bvh_leaf_ids, d.cap_leaf, (size_t)m);
    L.nodes = d.bvh; L.leaf_ids = d.bvh_leaf_ids;
    const int T = 256;
    lbvh_keys_kernel<<<(n_keys + T - 1) / T, T, 0, d.stream>>>(L);
    for (int k = 2; k <= n_keys; k *= 2)
        for (int j = k / 2; j >= 1; j /= 2) lbvh_bitonic_kernel<<<(n_keys + T - 1) / T, T, 0, d.stream>>>(L.keys, n_keys, j, k);
    if (m > 1) lbvh_tree_kernel<<<(m - 1 + T - 1) / T, T, 0, d.stream>>>(L);
    lbvh_boxes_kernel<<<(m + T - 1) / T, T, 0, d.stream>>>(L);
    CK(cudaGetLastError());
    int root_depth = 0;
    CK(cudaMemcpyAsync(&root_depth, L.depth, sizeof(int), cudaMemcpyDeviceToHost, d.stream));
    CK(cudaStreamSynchronize(d.stream));
    if (root_depth < 1 || root_depth > 40) return false;
    d.bvh_present = true;
    return true;
}

// ---- two-level cluster tables for cluster_scan (tray_kernels.cuh) ------------------------------------------------
// Spheres the pre-filter can bound are split recursively at the median of the longest centroid axis into words of 512, those
// into groups of 64 and those into chunks of 8 (every part full except the last; a word is padded to 8 group slots, a group to
// 8 chunk slots), so that chunk = 8 consecutive slots, group = 8 consecutive chunks and word = 8 consecutive groups are
// spatially compact; every level gets its conservative fp32 box. Spheres outside the filter's range (|C|inf > 256, r^2 > 256, non-finite) and,
// when there are only a few, spheres much larger than the rest go into "always" groups that every ray scans.
struct ClusterHost {
    std::vector<float4> blob;
    int off_box2 = 0, off_box1 = 0, off_box0 = 0, off_ids = 0, real_groups = 0, always_groups = 0, words8 = 0;
    unsigned always_last = 0;
    float r = 0;
    int unfilterable = 0;
};

void cluster_split(const tray_scene_desc* sc, std::vector<int>& ids, int first, int n, int leaf, std::vector<std::pair<int, int>>& out) {
    if (n <= leaf) { if (n > 0) out.push_back({first, n}); return; }
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int j = 0; j < n; j++) {
        const int i = ids[first + j];
        const double c[3] = {sc->cx[i], sc->cy[i], sc->cz[i]};
        for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], c[k]); hi[k] = std::max(hi[k], c[k]); }
    }
    int axis = 0;
    if (hi[1] - lo[1] > hi[axis] - lo[axis]) axis = 1;
    if (hi[2] - lo[2] > hi[axis] - lo[axis]) axis = 2;
    const double* key = axis == 0 ? sc->cx : (axis == 1 ? sc->cy : sc->cz);
    const int parts = (n + leaf - 1) / leaf, left = (parts / 2) * leaf;  // the left part is a whole number of leaves
    std::nth_element(ids.begin() + first, ids.begin() + first + left, ids.begin() + first + n,
                     [key](int x, int y) { return key[x] < key[y] || (key[x] == key[y] && x < y); });
    cluster_split(sc, ids, first, left, leaf, out);
    cluster_split(sc, ids, first + left, n - left, leaf, out);
}

float f32_up(double x) {  // smallest float >= x
    float f = (float)x;
    return (double)f < x ? std::nextafterf(f, std::numeric_limits<float>::infinity()) : f;
}

ClusterHost build_clusters(const tray_scene_desc* sc, int pad_id) {
    ClusterHost H;
    const int n = sc->n;
    const float finf = std::numeric_limits<float>::infinity(), qnan = std::numeric_limits<float>::quiet_NaN();
    std::vector<int> pool, always;
    std::vector<double> rs;
    auto filterable = [&](int i) {
        const double cm = std::max(std::fabs(sc->cx[i]), std::max(std::fabs(sc->cy[i]), std::fabs(sc->cz[i])));
        const double r2 = sc->radius[i] * sc->radius[i];
        return cm <= 256.0 && r2 <= 256.0;  // false for NaN
    };
    for (int i = 0; i < n; i++) {
        if (filterable(i)) { pool.push_back(i); rs.push_back(std::fabs(sc->radius[i])); } else always.push_back(i);
    }
    H.unfilterable = (int)always.size();
    if (!pool.empty()) {  // a handful of spheres much larger than the typical one: keep their boxes out of the tree
        std::vector<double> sorted = rs;
        std::nth_element(sorted.begin(), sorted.begin() + sorted.size() / 2, sorted.end());
        const double med = sorted[sorted.size() / 2];
        std::vector<int> big, rest;
        for (int i : pool) (std::fabs(sc->radius[i]) > 3.0 * med ? big : rest).push_back(i);
        if (!big.empty() && always.size() + big.size() <= 16) { always.insert(always.end(), big.begin(), big.end()); pool = rest; }
    }
    std::sort(always.begin(), always.end());
    // slot order: real groups (chunks of the pool, group by group, 8 group slots per word of 512 spheres), then the always-groups
    std::vector<std::pair<int, int>> words, groups, chunks;
    cluster_split(sc, pool, 0, (int)pool.size(), 512, words);
    std::vector<std::vector<int>> chunk_ids;   // [chunk] -> sphere ids (ascending); 8 chunk slots per group
    for (auto& w : words) {
        groups.clear();
        cluster_split(sc, pool, w.first, w.second, 64, groups);
        for (auto& g : groups) {
            chunks.clear();
            cluster_split(sc, pool, g.first, g.second, 8, chunks);
            for (auto& c : chunks) {
                std::vector<int> v(pool.begin() + c.first, pool.begin() + c.first + c.second);
                std::sort(v.begin(), v.end());
                chunk_ids.push_back(v);
            }
            while (chunk_ids.size() % 8) chunk_ids.push_back({});
        }
        while (chunk_ids.size() % 64) chunk_ids.push_back({});
    }
    H.real_groups = (int)words.size() * 8;
    H.words8 = ((int)words.size() + 7) / 8 * 8;
    H.always_groups = ((int)always.size() + 63) / 64;
    for (size_t k = 0; k < always.size(); k += 8)
        chunk_ids.push_back(std::vector<int>(always.begin() + k, always.begin() + std::min(always.size(), k + 8)));
    {
        const int last_chunks = ((int)always.size() - (H.always_groups - 1) * 64 + 7) / 8;  // chunks in the last always-group
        H.always_last = H.always_groups ? (0xffu << (8 - last_chunks)) & 0xffu : 0u;
    }
    while (chunk_ids.size() % 8) chunk_ids.push_back({});
    const int n_chunks = (int)chunk_ids.size(), n_groups_all = n_chunks / 8;
    // boxes in fp64: exact bounds of the member spheres, padded outwards by 2^-20 relative + a tiny absolute
    struct Box { double lo[3], hi[3]; bool empty = true, infinite = false; };
    auto grow_box = [&](Box& b, int i) {
        if (!filterable(i)) { b.infinite = true; b.empty = false; return; }
        const double c[3] = {sc->cx[i], sc->cy[i], sc->cz[i]}, r = std::fabs(sc->radius[i]);
        for (int k = 0; k < 3; k++) {
            const double pad = (std::fabs(c[k]) + r + 1.0) * 9.6e-7;
            const double l = c[k] - r - pad, h = c[k] + r + pad;
            b.lo[k] = b.empty ? l : std::min(b.lo[k], l);
            b.hi[k] = b.empty ? h : std::max(b.hi[k], h);
        }
        b.empty = false;
    };
    std::vector<Box> box2(n_chunks), box1(n_groups_all), box0(H.words8 + 1);
    for (int c = 0; c < n_chunks; c++)
        for (int i : chunk_ids[c]) { grow_box(box2[c], i); grow_box(box1[c / 8], i); if (c / 64 < (int)words.size()) grow_box(box0[c / 64], i); }
    float rmax = 0.f;
    auto emit = [&](const Box* b0, const Box* b1, float4* out) {  // one pair of boxes -> three float4 (centre, half extent)
        float c[2][3], e[2][3];
        const Box* bs[2] = {b0, b1};
        for (int j = 0; j < 2; j++)
            for (int k = 0; k < 3; k++) {
                const Box* b = bs[j];
                if (!b || b->empty) { c[j][k] = 0.f; e[j][k] = -finf; continue; }   // never hit
                if (b->infinite) { c[j][k] = 0.f; e[j][k] = finf; continue; }       // always hit
                c[j][k] = (float)(0.5 * (b->lo[k] + b->hi[k]));
                e[j][k] = f32_up(std::max(b->hi[k] - (double)c[j][k], (double)c[j][k] - b->lo[k]));
                rmax = std::max(rmax, f32_up((double)std::fabs(c[j][k]) + (double)e[j][k]));
            }
        out[0] = make_float4(c[0][0], c[1][0], c[0][1], c[1][1]);
        out[1] = make_float4(c[0][2], c[1][2], e[0][0], e[1][0]);
        out[2] = make_float4(e[0][1], e[1][1], e[0][2], e[1][2]);
    };
    // [pairs][box2: one box per chunk][box1: per group][box0: per word of 8 groups, padded to a multiple of 8 words][ids]
    const int f4_pairs = n_chunks * 8, f4_box2 = n_chunks / 2 * 3, f4_box1 = H.real_groups / 2 * 3, f4_box0 = H.words8 / 2 * 3,
              f4_ids = (n_chunks * 8 * 2 + 15) / 16;
    H.off_box2 = f4_pairs; H.off_box1 = H.off_box2 + f4_box2; H.off_box0 = H.off_box1 + f4_box1; H.off_ids = H.off_box0 + f4_box0;
    H.blob.assign((size_t)H.off_ids + f4_ids + 8, make_float4(0, 0, 0, 0));  // + slack: nothing reads past it, kept for safety
    uint16_t* ids16 = reinterpret_cast<uint16_t*>(H.blob.data() + H.off_ids);
    float mc = 0.f;
    for (int c = 0; c < n_chunks; c++) {
        float cc[8][3], nk[8];
        for (int u = 0; u < 8; u++) {
            const bool have = u < (int)chunk_ids[c].size();
            const int i = have ? chunk_ids[c][u] : -1;
            ids16[c * 8 + u] = (uint16_t)(have ? i : pad_id);
            if (!have) { cc[u][0] = cc[u][1] = cc[u][2] = 0.f; nk[u] = -finf; continue; }
            if (!filterable(i)) { cc[u][0] = cc[u][1] = cc[u][2] = qnan; nk[u] = qnan; continue; }
            const double r2 = sc->radius[i] * sc->radius[i];
            cc[u][0] = (float)sc->cx[i]; cc[u][1] = (float)sc->cy[i]; cc[u][2] = (float)sc->cz[i];
            nk[u] = -(float)((sc->cx[i] * sc->cx[i] + sc->cy[i] * sc->cy[i] + sc->cz[i] * sc->cz[i]) - r2);
            mc = std::max(mc, (float)std::max(std::fabs(sc->cx[i]), std::max(std::fabs(sc->cy[i]), std::fabs(sc->cz[i]))) * 1.0000002f);
        }
        for (int p = 0; p < 4; p++) {  // same pair layout as the linear table: (Cx0,Cx1,Cy0,Cy1), (Cz0,Cz1,-K0,-K1)
            H.blob[(size_t)c * 8 + 2 * p] = make_float4(cc[2 * p][0], cc[2 * p + 1][0], cc[2 * p][1], cc[2 * p + 1][1]);
            H.blob[(size_t)c * 8 + 2 * p + 1] = make_float4(cc[2 * p][2], cc[2 * p + 1][2], nk[2 * p], nk[2 * p + 1]);
        }
    }
    for (int c = 0; c < n_chunks; c += 2) emit(&box2[c], &box2[c + 1], H.blob.data() + H.off_box2 + c / 2 * 3);
    for (int g = 0; g < H.real_groups; g += 2) emit(&box1[g], &box1[g + 1], H.blob.data() + H.off_box1 + g / 2 * 3);  // (empty group slots: never hit)
    for (int w = 0; w < H.words8; w += 2) emit(&box0[w], &box0[w + 1], H.blob.data() + H.off_box0 + w / 2 * 3);
    H.r = std::max(rmax, mc) * 1.0000002f;
    return H;
}

constexpr int kClusterMaxSpheres = 32768;  // 64 words x 8 groups x 8 chunks x 8 spheres (the marked words of a segment are one 64-bit mask; slot ids are 16 bits)
constexpr int kBandRows = 1;  // rows per band of the tile split: single rows balance best (2160 rows over 8 shards: 270 each; 8-row bands gave 33 or 34 bands and 0.94 scaling efficiency at 8 GPUs)
// Scratch per pass, in samples (x24 bytes), whole pixels per pass: at least this much, more when the device has the room
// (tray_render sizes a pass from cudaMemGetInfo so that a device's whole share is one pass when it fits).
constexpr unsigned long long kPassSamples = 144ull << 20;

// The closest-hit BVH of the scene last uploaded (kept in ctx->bvh_*), on every device of the context: LBVH on the device for
// large scenes (or when asked for), the host's median-split build otherwise / as fallback.
void build_bvh_now(tray_ctx* ctx) {
    if (!ctx->bvh_pending) return;
    tray_scene_desc view{};
    view.n = (int32_t)ctx->bvh_r.size();
    view.cx = ctx->bvh_cx.data(); view.cy = ctx->bvh_cy.data(); view.cz = ctx->bvh_cz.data(); view.radius = ctx->bvh_r.data();
    const tray_scene_desc* sc = &view;
    std::vector<int> tree_ids;
    HostBvh hb = bvh_classify(sc, tree_ids);
    const bool want_device = ctx->bvh_build == TRAY_BVH_BUILD_DEVICE || (ctx->bvh_build == TRAY_BVH_BUILD_AUTO && tree_ids.size() >= 1024);
    bool host_built = false;
    if (!want_device) { bvh_build_host(sc, hb, tree_ids); host_built = true; }
    ctx->bvh_built_on_device = 0;
    for (Device& d : ctx->devs) {
        CK(cudaSetDevice(d.dev));
        CK(cudaStreamSynchronize(d.stream));
        d.bvh_present = false;
        d.bvh_n_always = (int)hb.always.size(); d.bvh_extent = hb.extent;
        if (!hb.always.empty()) {
            grow(d.bvh_always, d.cap_always, hb.always.size());
            CK(cudaMemcpyAsync(d.bvh_always, hb.always.data(), sizeof(int) * hb.always.size(), cudaMemcpyHostToDevice, d.stream));
        }
        bool on_device = false;
        if (want_device && !tree_ids.empty()) on_device = bvh_build_device(d, sc, tree_ids);
        if (on_device) ctx->bvh_built_on_device = 1;
        else {
            if (!host_built) { std::vector<int> tmp = tree_ids; bvh_build_host(sc, hb, tmp); host_built = true; }
            if (!hb.nodes.empty()) {
                grow(d.bvh, d.cap_bvh, hb.nodes.size());
                CK(cudaMemcpyAsync(d.bvh, hb.nodes.data(), sizeof(BvhNode) * hb.nodes.size(), cudaMemcpyHostToDevice, d.stream));
                grow(d.bvh_leaf_ids, d.cap_leaf, hb.leaf_ids.size());
                d.bvh_present = true;
                CK(cudaMemcpyAsync(d.bvh_leaf_ids, hb.leaf_ids.data(), sizeof(int) * hb.leaf_ids.size(), cudaMemcpyHostToDevice, d.stream));
            }
        }
        CK(cudaStreamSynchronize(d.stream));
    }
    ctx->bvh_pending = false;
}

}  // namespace

extern "C" {

int tray_abi_version(void) { return TRAY_ABI_VERSION; }

const char* tray_last_error(tray_ctx* ctx) { return ctx ? ctx->err.c_str() : g_init_error.c_str(); }

int tray_init(const int* devices, int n_devices, tray_ctx** out) {
    if (!out) return fail(nullptr, TRAY_E_INVALID, "tray_init: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, TRAY_E_NO_DEVICE, std::string("tray_init: no CUDA device (") + cudaGetErrorString(e) + "); there is no CPU fallback");
    std::vector<int> ids;
    if (!devices || n_devices <= 0) ids.push_back(0);
    else ids.assign(devices, devices + n_devices);
    if (ids.size() > 8) return fail(nullptr, TRAY_E_INVALID, "tray_init: at most 8 devices per context");
    tray_ctx* ctx = new tray_ctx();
    try {
        double thr[256];
        build_srgb_thresholds(thr);
        for (int id : ids) {
            if (id < 0 || id >= count) throw std::runtime_error("tray_init: device index out of range");
            cudaDeviceProp prop;
            CK(cudaGetDeviceProperties(&prop, id));
            if (prop.major < 10) throw std::runtime_error(std::string("tray_init: device ") + prop.name + " is not sm_100 class; libtraycuda is built for sm_100a only");
            Device d;
            d.dev = id;
            d.num_sms = prop.multiProcessorCount;
            CK(cudaSetDevice(id));
            CK(cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking));
            CK(cudaEventCreate(&d.ev_begin)); CK(cudaEventCreate(&d.ev_end));
            CK(cudaMalloc(&d.stats, 8 * sizeof(unsigned long long)));
            CK(cudaMemset(d.stats, 0, 8 * sizeof(unsigned long long)));
            CK(cudaMalloc(&d.srgb_thr, 256 * sizeof(double)));
            CK(cudaMemcpy(d.srgb_thr, thr, sizeof thr, cudaMemcpyHostToDevice));
            CK(cudaMemcpyToSymbol(g_zig_kn, ZIG_KN_INIT, sizeof(uint32_t) * 128));
            CK(cudaMemcpyToSymbol(g_zig_wn, ZIG_WN_INIT, sizeof(float) * 128));
            CK(cudaMemcpyToSymbol(g_zig_fn, ZIG_FN_INIT, sizeof(float) * 128));
            ctx->devs.push_back(d);
        }
        // peer access for the sample-split combine (root reads peers' partial sums over NVLink); without it the sums are staged
        for (size_t i = 1; i < ctx->devs.size(); i++) {
            int can = 0;
            CK(cudaDeviceCanAccessPeer(&can, ctx->devs[0].dev, ctx->devs[i].dev));
            ctx->devs[i].peer_of_root = false;
            if (can) {
                CK(cudaSetDevice(ctx->devs[0].dev));
                cudaError_t pe = cudaDeviceEnablePeerAccess(ctx->devs[i].dev, 0);
                if (pe == cudaSuccess || pe == cudaErrorPeerAccessAlreadyEnabled) ctx->devs[i].peer_of_root = true;
                cudaGetLastError();
            }
        }
        for (Device& d : ctx->devs) {
            CK(cudaSetDevice(d.dev));
            CK(cudaStreamCreateWithFlags(&d.peek_stream, cudaStreamNonBlocking));
            CK(cudaHostAlloc(&d.peek_host, sizeof(unsigned long long), cudaHostAllocDefault));
        }
        CK(cudaSetDevice(ctx->devs[0].dev));
    } catch (const std::exception& ex) {
        g_init_error = ex.what();
        tray_destroy(ctx);
        return TRAY_E_CUDA;
    }
    *out = ctx;
    return TRAY_OK;
}

void tray_destroy(tray_ctx* ctx) {
    if (!ctx) return;
    for (Device& d : ctx->devs) {
        cudaSetDevice(d.dev);
        cudaStreamSynchronize(d.stream);
        free_scene(d);
        cudaFree(d.scratch); cudaFree(d.rgba); cudaFree(d.hdr); cudaFree(d.counters); cudaFree(d.stk_g); cudaFree(d.gen);
        cudaFree(d.lbvh_keys); cudaFree(d.lbvh_int); cudaFree(d.lbvh_int2);
        cudaFree(d.wf_rec); cudaFree(d.wf_queue); cudaFree(d.wf_free); cudaFree(d.wf_cnt);
        if (d.wf_host) cudaFreeHost(d.wf_host);
        cudaFree(d.pr_tmp); cudaFree(d.pr_small); cudaFree(d.pr_ansi);
        cudaFree(d.png_filt); cudaFree(d.png_out); cudaFree(d.png_hist); cudaFree(d.png_piece); cudaFree(d.png_adler); cudaFree(d.png_blocks); cudaFree(d.png_tot); cudaFree(d.stats); cudaFree(d.srgb_thr);
        if (d.pinned) cudaFreeHost(d.pinned);
        if (d.ev_begin) cudaEventDestroy(d.ev_begin);
        for (cudaEvent_t e : d.ev_pool) cudaEventDestroy(e);
        if (d.ev_end) cudaEventDestroy(d.ev_end);
        if (d.stream) cudaStreamDestroy(d.stream);
        if (d.peek_stream) cudaStreamDestroy(d.peek_stream);
        if (d.peek_host) cudaFreeHost(d.peek_host);
    }
    for (Device& d : ctx->devs)
        if (d.stage) { cudaSetDevice(ctx->devs[0].dev); cudaFree(d.stage); }
    delete ctx;
}

int tray_scene_upload(tray_ctx* ctx, const tray_scene_desc* sc) {
    if (!ctx) return TRAY_E_INVALID;
    std::lock_guard<std::mutex> lock(ctx->mu);
    if (!sc || sc->n < 0 || (sc->n > 0 && (!sc->cx || !sc->cy || !sc->cz || !sc->radius || !sc->mat_kind || !sc->mat_params)))
        return fail(ctx, TRAY_E_INVALID, "tray_scene_upload: null array");
    if (sc->n > 65535) return fail(ctx, TRAY_E_UNSUPPORTED, "tray_scene_upload: more than 65535 spheres");
    for (int i = 0; i < sc->n; i++)
        if (sc->mat_kind[i] > TRAY_MAT_DIELECTRIC) return fail(ctx, TRAY_E_UNSUPPORTED, "tray_scene_upload: unknown material kind (only Lambertian/Metal/Dielectric spheres are supported; no CPU fallback)");
    try {
        const int ch = TRAY_CH > 8 ? TRAY_CH : 8;
        int n = sc->n, n_pad = std::max(ch, (n + ch - 1) / ch * ch);  // the hot loop consumes chunks of TRAY_CH spheres
        const double ninf = -std::numeric_limits<double>::infinity();
        const int n_alloc = n_pad + 8;  // entry n is always a never-hit padding entry (the cluster tables point their empty slots at it)
        std::vector<double4> gd(n_alloc); std::vector<float4> gf(n_alloc);
        std::vector<double> rd(n_alloc, 1.0); std::vector<float> rf(n_alloc, 1.0f);
        std::vector<uint8_t> kd(n_alloc, 0); std::vector<double4> pr(n_alloc);
        for (int i = 0; i < n_alloc; i++) {
            if (i < n) {
                double r = sc->radius[i];
                gd[i] = make_double4(sc->cx[i], sc->cy[i], sc->cz[i], r * r);  // Radius*Radius, objects.go:85
                float fr = (float)r;
                gf[i] = make_float4((float)sc->cx[i], (float)sc->cy[i], (float)sc->cz[i], fr * fr);
                rd[i] = r; rf[i] = fr; kd[i] = sc->mat_kind[i];
                pr[i] = make_double4(sc->mat_params[4 * i], sc->mat_params[4 * i + 1], sc->mat_params[4 * i + 2], sc->mat_params[4 * i + 3]);
            } else {  // padding: c = +inf, disc = -inf -> certainly missed
                gd[i] = make_double4(0, 0, 0, ninf);
                gf[i] = make_float4(0, 0, 0, (float)ninf);
                pr[i] = make_double4(0, 0, 0, 0);
            }
        }
        // fp32 pre-filter table: pairs of spheres, two float4 each: (Cx0,Cx1,Cy0,Cy1), (Cz0,Cz1,-K0,-K1) with K = |C|^2 - r^2
        // (fp64 on the host, then rounded). Spheres whose magnitudes would blow up the filter's error bound
        // (|c|_inf > 256 or r*r > 256, e.g. the r=1000 ground sphere) are stored as NaN: never "certainly missed", i.e.
        // always exact-tested. Padding: -K = -inf (always missed).
        std::vector<float4> fp((size_t)n_pad);
        float mc = 0.f, r2max = 0.f;
        const float qnan = std::numeric_limits<float>::quiet_NaN(), finf = std::numeric_limits<float>::infinity();
        for (int j = 0; j < n_pad / 2; j++) {
            float c[2][3], nk[2];
            for (int k = 0; k < 2; k++) {
                int i = 2 * j + k;
                if (i >= n) { c[k][0] = c[k][1] = c[k][2] = 0.f; nk[k] = -finf; continue; }
                double cm = std::max(std::fabs(sc->cx[i]), std::max(std::fabs(sc->cy[i]), std::fabs(sc->cz[i])));
                double r2 = sc->radius[i] * sc->radius[i];
                if (!(cm <= 256.0) || !(r2 <= 256.0)) { c[k][0] = c[k][1] = c[k][2] = qnan; nk[k] = qnan; continue; }
                c[k][0] = (float)sc->cx[i]; c[k][1] = (float)sc->cy[i]; c[k][2] = (float)sc->cz[i];
                nk[k] = -(float)((sc->cx[i] * sc->cx[i] + sc->cy[i] * sc->cy[i] + sc->cz[i] * sc->cz[i]) - r2);
                mc = std::max(mc, (float)cm * 1.0000002f);
                r2max = std::max(r2max, (float)r2 * 1.0000002f);
            }
            fp[2 * j] = make_float4(c[0][0], c[1][0], c[0][1], c[1][1]);
            fp[2 * j + 1] = make_float4(c[0][2], c[1][2], nk[0], nk[1]);
        }
        // cluster tables (the default closest-hit structure up to kClusterMaxSpheres spheres: staged into shared memory and walked
        // on two levels up to kSmemClusterMax spheres, read from global memory and walked on three levels above)
        ClusterHost clh;
        const bool use_clusters = n <= kClusterMaxSpheres;
        if (use_clusters) clh = build_clusters(sc, n);
        ctx->cluster_unfilterable = use_clusters ? clh.unfilterable : n;
        ctx->bvh_built_on_device = 0;
        for (Device& d : ctx->devs) {
            CK(cudaSetDevice(d.dev));
            CK(cudaStreamSynchronize(d.stream));
            d.bvh_present = false;
            d.n = n; d.n_pad = n_pad;
            d.filt_mc = mc; d.filt_r2max = r2max;
            d.bvh_n_always = 0; d.bvh_extent = 0;
            grow(d.fpair, d.cap_fpair, (size_t)n_pad);
            CK(cudaMemcpyAsync(d.fpair, fp.data(), sizeof(float4) * n_pad, cudaMemcpyHostToDevice, d.stream));
            grow(d.geo_d, d.cap_geo_d, (size_t)n_alloc); grow(d.geo_f, d.cap_geo_f, (size_t)n_alloc);
            grow(d.radius_d, d.cap_radius_d, (size_t)n_alloc); grow(d.radius_f, d.cap_radius_f, (size_t)n_alloc);
            grow(d.kind, d.cap_kind, (size_t)n_alloc); grow(d.params, d.cap_params, (size_t)n_alloc);
            CK(cudaMemcpyAsync(d.geo_d, gd.data(), sizeof(double4) * n_alloc, cudaMemcpyHostToDevice, d.stream));
            CK(cudaMemcpyAsync(d.geo_f, gf.data(), sizeof(float4) * n_alloc, cudaMemcpyHostToDevice, d.stream));
            CK(cudaMemcpyAsync(d.radius_d, rd.data(), sizeof(double) * n_alloc, cudaMemcpyHostToDevice, d.stream));
            CK(cudaMemcpyAsync(d.radius_f, rf.data(), sizeof(float) * n_alloc, cudaMemcpyHostToDevice, d.stream));
            CK(cudaMemcpyAsync(d.kind, kd.data(), n_alloc, cudaMemcpyHostToDevice, d.stream));
            CK(cudaMemcpyAsync(d.params, pr.data(), sizeof(double4) * n_alloc, cudaMemcpyHostToDevice, d.stream));
            d.cl_present = false;
            if (use_clusters) {
                grow(d.cl_blob, d.cap_cl_blob, clh.blob.size());
                CK(cudaMemcpyAsync(d.cl_blob, clh.blob.data(), sizeof(float4) * clh.blob.size(), cudaMemcpyHostToDevice, d.stream));
                d.cl_blob_f4 = (int)clh.blob.size(); d.cl_off_box2 = clh.off_box2; d.cl_off_box1 = clh.off_box1; d.cl_off_box0 = clh.off_box0; d.cl_off_ids = clh.off_ids;
                d.cl_real_groups = clh.real_groups; d.cl_always_groups = clh.always_groups; d.cl_always_last = clh.always_last; d.cl_r = clh.r;
                d.cl_present = true;
            }
            // the tables travel as asynchronous copies on the device's stream (the pageable sources above are staged by the driver when the
            // call is made); one wait per device instead of one per table
            CK(cudaStreamSynchronize(d.stream));
        }
        // BVH: now if the scene needs it whatever the caller asks for (the cluster walk cannot serve it), on first use otherwise
        ctx->bvh_cx.assign(sc->cx, sc->cx + n); ctx->bvh_cy.assign(sc->cy, sc->cy + n); ctx->bvh_cz.assign(sc->cz, sc->cz + n);
        ctx->bvh_r.assign(sc->radius, sc->radius + n);
        ctx->bvh_pending = true;
        if (!(use_clusters && ctx->cluster_unfilterable <= kAutoClusterUnfilterable)) build_bvh_now(ctx);
        for (int i = 0; i < 3; i++) { ctx->bg_a[i] = sc->bg_a[i]; ctx->bg_b[i] = sc->bg_b[i]; }
        ctx->host_geo_d = gd; ctx->host_geo_f = gf;
        ctx->have_scene = true;
    } catch (const std::exception& ex) {
        return fail(ctx, TRAY_E_CUDA, ex.what());
    }
    return TRAY_OK;
}

static void copy_out(tray_ctx* ctx, uint8_t* rgba_out, size_t stride) {
    // gather: every device's local rows -> pinned staging -> caller rows
    const size_t row_bytes = (size_t)ctx->width * 4;
    for (Device& d : ctx->devs) {
        size_t nrows = d.local_rows.size();
        if (!nrows) continue;
        CK(cudaSetDevice(d.dev));
        size_t bytes = nrows * row_bytes;
        if (bytes > d.pinned_cap) {
            if (d.pinned) CK(cudaFreeHost(d.pinned));
            d.pinned = nullptr; d.pinned_cap = 0;
            CK(cudaHostAlloc(&d.pinned, bytes, cudaHostAllocDefault));
            d.pinned_cap = bytes;
        }
        CK(cudaMemcpyAsync(d.pinned, d.rgba, bytes, cudaMemcpyDeviceToHost, d.stream));
    }
    for (Device& d : ctx->devs) {
        size_t nrows = d.local_rows.size();
        if (!nrows) continue;
        CK(cudaSetDevice(d.dev));
        CK(cudaStreamSynchronize(d.stream));
        size_t r = 0;
        while (r < nrows) {  // coalesce runs of consecutive rows when the caller's stride is tight
            size_t e = r + 1;
            if (stride == row_bytes)
                while (e < nrows && d.local_rows[e] == d.local_rows[e - 1] + 1) e++;
            memcpy(rgba_out + (size_t)d.local_rows[r] * stride, d.pinned + r * row_bytes, (e - r) * row_bytes);
            r = e;
        }
    }
}

int tray_render(tray_ctx* ctx, const tray_camera* cam, const tray_params* p, uint8_t* rgba_out, size_t stride, tray_stats* stats) {
    if (!ctx) return TRAY_E_INVALID;
    std::lock_guard<std::mutex> lock(ctx->mu);
    if (!cam || !p) return fail(ctx, TRAY_E_INVALID, "tray_render: null camera/params");
    if (!ctx->have_scene) return fail(ctx, TRAY_E_NO_SCENE, "tray_render: no scene uploaded");
    if (p->width <= 0 || p->height <= 0 || p->spp <= 0 || p->max_depth < 0)
        return fail(ctx, TRAY_E_INVALID, "tray_render: width/height/spp must be > 0 and max_depth >= 0 (apply Tracer defaults on the host)");
    if (p->stream_mode != TRAY_STREAM_REFERENCE && p->stream_mode != TRAY_STREAM_PER_SAMPLE) return fail(ctx, TRAY_E_INVALID, "tray_render: bad stream_mode");
    if (p->max_depth > kMaxDepth) return fail(ctx, TRAY_E_UNSUPPORTED, "tray_render: max_depth > 256");
    if (p->y0 < 0 || p->y1 > p->height || p->y0 > p->y1) return fail(ctx, TRAY_E_INVALID, "tray_render: bad row range");
    if (rgba_out && stride < (size_t)p->width * 4) return fail(ctx, TRAY_E_INVALID, "tray_render: stride < 4*width");
    if (p->precision < TRAY_FP64_FMA || p->precision > TRAY_FP64_STRICT_BRUTE) return fail(ctx, TRAY_E_INVALID, "tray_render: bad precision");
    if (p->accel < TRAY_ACCEL_AUTO || p->accel > TRAY_ACCEL_CLUSTER) return fail(ctx, TRAY_E_INVALID, "tray_render: bad accel");
    if (p->layout < TRAY_LAYOUT_AUTO || p->layout > TRAY_LAYOUT_WAVEFRONT) return fail(ctx, TRAY_E_INVALID, "tray_render: bad layout");
    if (p->seed == 0) return fail(ctx, TRAY_E_INVALID, "tray_render: seed 0 (the host shim must draw a random seed, ray/tracer.go:32)");
    auto t_start = std::chrono::steady_clock::now();
    if (p->sums_mode < TRAY_SUMS_OFF || p->sums_mode > TRAY_SUMS_ACCUMULATE) return fail(ctx, TRAY_E_INVALID, "tray_render: bad sums_mode");
    if (p->sums_mode != TRAY_SUMS_OFF && p->stream_mode != TRAY_STREAM_PER_SAMPLE)
        return fail(ctx, TRAY_E_INVALID, "tray_render: sample subsets need per-sample streams");
    const bool subset = p->sums_mode != TRAY_SUMS_OFF;
    const int rows = p->y1 - p->y0;
    const int G = (int)ctx->devs.size();
    int launches = 0, trace_launch_count = 0;
    try {
        ctx->width = p->width; ctx->height = p->height; ctx->y0 = p->y0; ctx->y1 = p->y1;
        ctx->have_image = false; ctx->have_hdr = false;
        ctx->progress_base = 0; ctx->progress_spp = p->spp; ctx->rendering = true;
        for (Device& d : ctx->devs) { d.local_rows.clear(); d.passes = 0; d.ev_used = 0; d.timed = false; }
        DevCamera dcam = dev_camera(cam, ctx->indisc_variant);
        if (ctx->bvh_pending && p->stream_mode != TRAY_STREAM_REFERENCE && closest_hit_structure(ctx, ctx->devs[0], p->accel) == kUseBvh) build_bvh_now(ctx);
        const int ext_count = p->shard_count > 1 ? p->shard_count : 1;
        const int ext_index = p->shard_count > 1 ? p->shard_index : 0;
        if (ext_index < 0 || ext_index >= ext_count) throw std::runtime_error("tray_render: shard_index out of range");

        if (p->stream_mode == TRAY_STREAM_REFERENCE) {
            // ---- conformance: reference chunk streams on device 0, one warp per chunk ----
            if (p->precision == TRAY_FP32) throw std::runtime_error("tray_render: reference streams are fp64 only");
            if (ctx->devs[0].n_pad * sizeof(double4) + sizeof(ZigTables) > kSmemBudget) throw std::runtime_error("tray_render: scene too large for the conformance kernel");
            Device& d = ctx->devs[0];
            CK(cudaSetDevice(d.dev));
            RefArgs A;
            A.cam = dcam; A.width = p->width; A.spp = p->spp; A.max_depth = p->max_depth; A.ray_radius = p->ray_radius;
            A.seed = p->seed; A.row_begin = p->y0; A.row_end = p->y1;
            if (p->num_workers <= 0 || p->stream_idx >= 0) {  // RenderLines(idx, y0, y1): one stream
                A.chunk_rows = std::max(rows, 1); A.n_chunks = 1; A.idx_override = p->stream_idx >= 0 ? p->stream_idx : p->y0;
            } else if (p->num_workers == 1) {                   // tracer.go:87-89
                A.chunk_rows = std::max(rows, 1); A.n_chunks = 1; A.idx_override = 0;
            } else {                                            // tracer.go:93-103
                int chunk = std::max(4, p->height / (p->num_workers * 4));
                A.chunk_rows = chunk; A.n_chunks = (rows + chunk - 1) / chunk; A.idx_override = -1;
            }
            size_t px = (size_t)rows * p->width;
            grow(d.rgba, d.rgba_cap, std::max<size_t>(px * 4, 4));
            grow(d.hdr, d.hdr_cap, std::max<size_t>(px * 3, 3));
            A.rgba = d.rgba; A.hdr = d.hdr; A.srgb_thr = d.srgb_thr; A.stats = d.stats;
            CK(cudaMemsetAsync(d.stats, 0, 8 * sizeof(unsigned long long), d.stream));
            CK(cudaEventRecord(d.ev_begin, d.stream));
            d.timed = true;
            if (rows > 0 && p->max_depth == 0) {
                black_kernel<<<(unsigned)((px + 255) / 256), 256, 0, d.stream>>>(d.rgba, d.hdr, px);
                CK(cudaGetLastError());
                launches++;
            } else if (rows > 0) {
                CK(cudaEventRecord(next_event(d), d.stream));
                size_t smem = (size_t)d.n_pad * sizeof(double4) + sizeof(ZigTables);
                DevScene<double> S = dev_scene<double>(ctx, d);
                if (p->precision == TRAY_FP64_FMA) {
                    CK(cudaFuncSetAttribute(reference_stream_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    reference_stream_kernel<true><<<A.n_chunks, 32, smem, d.stream>>>(A, S);
                } else {
                    CK(cudaFuncSetAttribute(reference_stream_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    reference_stream_kernel<false><<<A.n_chunks, 32, smem, d.stream>>>(A, S);
                }
                CK(cudaGetLastError());
                CK(cudaEventRecord(next_event(d), d.stream));
                launches++;
                trace_launch_count++;
            }
            CK(cudaEventRecord(d.ev_end, d.stream));
            for (int r = 0; r < rows; r++) d.local_rows.push_back(p->y0 + r);
        } else {
            // ---- throughput mode: per-sample streams, persistent megakernel ----
            const bool split_samples = (p->split_mode == TRAY_SPLIT_SAMPLES) && G > 1;
            if (split_samples && ext_count > 1) throw std::runtime_error("tray_render: sample split cannot be combined with external shards");
            if (subset) {
                if (p->max_depth == 0) throw std::runtime_error("tray_render: sums_mode needs max_depth > 0");
                if (split_samples) throw std::runtime_error("tray_render: sums_mode cannot be combined with the in-context sample split");
                if (p->sample_stride < 1 || p->sample_count < 1 || p->sample_offset < 0 ||
                    (long long)p->sample_offset + (long long)(p->sample_count - 1) * p->sample_stride >= (long long)p->spp)
                    throw std::runtime_error("tray_render: sample subset out of range (need offset + (count-1)*stride < spp)");
                const int key[6] = {p->width, p->height, p->y0, p->y1, ext_index, ext_count};
                if (p->sums_mode == TRAY_SUMS_ACCUMULATE) {
                    if (!ctx->hdr_is_sums || memcmp(key, ctx->sums_key, sizeof key) != 0)
                        throw std::runtime_error("tray_render: TRAY_SUMS_ACCUMULATE needs sums of the same geometry from a previous TRAY_SUMS_* render");
                } else {
                    ctx->sums_samples = 0;
                }
                memcpy(ctx->sums_key, key, sizeof key);
            }
            for (int g = 0; g < G; g++) {
                Device& d = ctx->devs[g];
                CK(cudaSetDevice(d.dev));
                // which rows does this device own?
                int shard_count, shard_index;
                if (split_samples) { shard_count = ext_count; shard_index = ext_index; }
                else { shard_count = ext_count * G; shard_index = ext_index * G + g; }
                for (int r = 0; r < rows; r++)
                    if (shard_count <= 1 || (r / kBandRows) % shard_count == shard_index) d.local_rows.push_back(p->y0 + r);
                const unsigned long long n_pixels = (unsigned long long)d.local_rows.size() * p->width;
                int spp_local = p->spp, stride_s = 1, offset_s = 0;
                if (split_samples) {
                    stride_s = G; offset_s = g;
                    spp_local = p->spp > g ? (p->spp - g + G - 1) / G : 0;
                }
                if (subset) { stride_s = p->sample_stride; offset_s = p->sample_offset; spp_local = p->sample_count; }
                CK(cudaMemsetAsync(d.stats, 0, 8 * sizeof(unsigned long long), d.stream));
                grow(d.rgba, d.rgba_cap, std::max<size_t>(n_pixels * 4, 4));
                grow(d.hdr, d.hdr_cap, std::max<size_t>(n_pixels * 3, 3));
                CK(cudaEventRecord(d.ev_begin, d.stream));
                d.timed = true;
                if (n_pixels == 0 || spp_local == 0) {
                    if (split_samples && n_pixels) CK(cudaMemsetAsync(d.hdr, 0, n_pixels * 3 * sizeof(double), d.stream));
                    CK(cudaEventRecord(d.ev_end, d.stream));
                    continue;
                }
                if (p->max_depth == 0) {  // every sample is black; in the sample split the combine kernel turns the zero sums into pixels
                    black_kernel<<<(unsigned)((n_pixels + 255) / 256), 256, 0, d.stream>>>(d.rgba, d.hdr, n_pixels);
                    CK(cudaGetLastError());
                    launches++;
                    CK(cudaEventRecord(d.ev_end, d.stream));
                    continue;
                }
                // One pass per device whenever its scratch (24 B per sample) fits in a share of the free memory: a pass ends with a
                // tail in which the SMs run dry one by one, so two passes cost two tails (config 3 on 8 GPUs: 265 M samples per rank).
                unsigned long long pass_samples = kPassSamples;
                const unsigned long long share = n_pixels * (unsigned long long)spp_local;
                if (share > pass_samples) {
                    if (share * 3ull <= (unsigned long long)d.scratch_cap && share <= 0xfff00000ull) {
                        pass_samples = share;  // the scratch of an earlier render already holds the whole share: no driver query
                    } else {
                        size_t free_b = 0, total_b = 0;
                        CK(cudaMemGetInfo(&free_b, &total_b));  // (milliseconds on a 180 GB device: only when the scratch has to grow)
                        const unsigned long long fit = (unsigned long long)(0.45 * (double)(free_b + d.scratch_cap * sizeof(double))) / 24ull;
                        pass_samples = std::max(pass_samples, std::min(fit, 0xfff00000ull));  // (sample indices of a pass are 32-bit)
                    }
                }
                unsigned long long px_per_pass = std::max<unsigned long long>(1, pass_samples / (unsigned)spp_local);
                px_per_pass = std::min(px_per_pass, n_pixels);
                int n_pass = (int)((n_pixels + px_per_pass - 1) / px_per_pass);
                grow(d.scratch, d.scratch_cap, (size_t)(px_per_pass * spp_local * 3));
                if (n_pass > d.counters_cap) {
                    if (d.counters) CK(cudaFree(d.counters));
                    d.counters = nullptr;
                    CK(cudaMalloc(&d.counters, sizeof(unsigned long long) * n_pass));
                    d.counters_cap = n_pass;
                }
                CK(cudaEventRecord(d.ev_begin, d.stream));  // (again: the device time of a render starts after the host-side sizing above)
                CK(cudaMemsetAsync(d.counters, 0, sizeof(unsigned long long) * n_pass, d.stream));
                d.passes = n_pass;
                for (int ps = 0; ps < n_pass; ps++) {
                    unsigned long long p0 = (unsigned long long)ps * px_per_pass;
                    unsigned long long npx = std::min(px_per_pass, n_pixels - p0);
                    TraceArgs A;
                    A.cam = dcam; A.width = p->width; A.spp = p->spp; A.max_depth = p->max_depth;
                    A.ray_radius = p->ray_radius; A.seed = p->seed;
                    A.n_samples = npx * spp_local; A.pass_pixel0 = p0; A.row0 = p->y0;
                    A.band_rows = kBandRows; A.shard_count = shard_count; A.shard_index = shard_index;
                    A.spp_local = spp_local; A.sample_stride = stride_s; A.sample_offset = offset_s;
                    A.counter = d.counters + ps; A.scratch = d.scratch; A.stats = d.stats; A.progress = d.stats + 2;
                    A.stk_g = nullptr; A.n_slots = 0; A.gen = nullptr;
                    const bool wavefront = wavefront_runs(ctx, d, p->precision, p->accel, p->layout);  // generates its rays in wf_regen
                    if (!wavefront && !TRAY_GEN_INKERNEL) {
                        grow(d.gen, d.gen_cap, (size_t)A.n_samples);
                        camera_ray_kernel<<<(unsigned)((A.n_samples + 255) / 256), 256, 0, d.stream>>>(A, d.gen);
                        CK(cudaGetLastError());
                        A.gen = d.gen;
                        launches++;
                    }
                    CK(cudaEventRecord(next_event(d), d.stream));
                    const int n_tr = launch_trace(ctx, d, A, p->precision, p->accel, p->layout);
                    launches += n_tr - 1;
                    trace_launch_count += n_tr;
                    CK(cudaEventRecord(next_event(d), d.stream));
                    ResolveArgs R;
                    R.scratch = d.scratch; R.n_pixels = npx; R.pass_pixel0 = p0; R.spp_local = spp_local;
                    R.inv_spp = 1.0 / (double)p->spp; R.partial = subset ? p->sums_mode : (split_samples ? 1 : 0);
                    R.rgba = d.rgba; R.hdr = d.hdr; R.srgb_thr = d.srgb_thr;
                    resolve_kernel<<<(unsigned)((3 * npx + 255) / 256), 256, 0, d.stream>>>(R);
                    CK(cudaGetLastError());
                    launches += 2;
                }
                CK(cudaEventRecord(d.ev_end, d.stream));
            }
            if (split_samples) {
                // root sums the partial HDR sums of all devices through peer pointers, then sRGB: one kernel
                for (int g = 1; g < G; g++) { CK(cudaSetDevice(ctx->devs[g].dev)); CK(cudaStreamSynchronize(ctx->devs[g].stream)); }
                Device& d0 = ctx->devs[0];
                CK(cudaSetDevice(d0.dev));
                const unsigned long long n_pixels = (unsigned long long)d0.local_rows.size() * p->width;
                CombineArgs C;
                for (int g = 0; g < G; g++) {
                    Device& dg = ctx->devs[g];
                    C.partial[g] = dg.hdr;
                    if (g > 0 && !dg.peer_of_root && n_pixels) {  // no peer mapping: bring the partial sums over with a peer copy
                        grow(dg.stage, dg.stage_cap, (size_t)n_pixels * 3);
                        CK(cudaMemcpyPeerAsync(dg.stage, d0.dev, dg.hdr, dg.dev, n_pixels * 3 * sizeof(double), d0.stream));
                        C.partial[g] = dg.stage;
                    }
                }
                C.n_parts = G; C.n_pixels = n_pixels; C.inv_spp = 1.0 / (double)p->spp;
                C.rgba = d0.rgba; C.srgb_thr = d0.srgb_thr;
                // combine in place is not possible (hdr is an input): reuse scratch for the mean
                grow(d0.scratch, d0.scratch_cap, std::max<size_t>(n_pixels * 3, 3));
                C.hdr = d0.scratch;
                if (n_pixels) {
                    combine_kernel<<<(unsigned)((n_pixels + 255) / 256), 256, 0, d0.stream>>>(C);
                    CK(cudaGetLastError());
                    CK(cudaMemcpyAsync(d0.hdr, d0.scratch, n_pixels * 3 * sizeof(double), cudaMemcpyDeviceToDevice, d0.stream));
                    launches++;
                }
                CK(cudaEventRecord(d0.ev_end, d0.stream));
                for (int g = 1; g < G; g++) ctx->devs[g].local_rows.clear();  // only the root holds the image
            }
        }
        ctx->split_mode = p->split_mode;
        if (rgba_out && !subset) copy_out(ctx, rgba_out, stride);
        double kernel_ms = 0, trace_ms = 0;
        unsigned long long seg = 0, exh = 0, bvh_tests = 0, box_tests = 0, violations = 0;
        for (Device& d : ctx->devs) {
            CK(cudaSetDevice(d.dev));
            CK(cudaStreamSynchronize(d.stream));
            if (!d.timed) continue;  // (reference streams run on device 0 only: the other devices of the context did nothing)
            float ms = 0, ms2 = 0;
            CK(cudaEventElapsedTime(&ms, d.ev_begin, d.ev_end));
            for (int e = 0; e + 1 < d.ev_used; e += 2) {
                float t = 0;
                CK(cudaEventElapsedTime(&t, d.ev_pool[e], d.ev_pool[e + 1]));
                ms2 += t;
            }
            kernel_ms = std::max(kernel_ms, (double)ms);
            trace_ms = std::max(trace_ms, (double)ms2);
            unsigned long long st[5];
            CK(cudaMemcpy(st, d.stats, sizeof st, cudaMemcpyDeviceToHost));
            seg += st[0]; exh += st[1]; bvh_tests += st[3]; box_tests += st[4];
#ifdef TRAY_BOUNDS_CHECK
            unsigned long long gv = 0;  // running total of this device in this process
            CK(cudaMemcpyFromSymbol(&gv, g_bounds_violations, sizeof gv));
            violations += gv;
#endif
        }
        ctx->have_image = !subset; ctx->have_hdr = true;
        ctx->hdr_is_sums = subset;
        if (subset) ctx->sums_samples += (uint64_t)p->sample_count;
        ctx->rendering = false;
        unsigned long long my_rows = 0;
        if (p->stream_mode == TRAY_STREAM_REFERENCE || (p->split_mode == TRAY_SPLIT_SAMPLES && G > 1)) {
            for (int r = 0; r < rows; r++)
                if (ext_count <= 1 || (r / kBandRows) % ext_count == ext_index || p->stream_mode == TRAY_STREAM_REFERENCE) my_rows++;
        } else {
            for (Device& d : ctx->devs) my_rows += d.local_rows.size();
        }
        ctx->progress_base = my_rows * (unsigned long long)p->width;
        if (stats) {
            memset(stats, 0, sizeof *stats);
            stats->paths = my_rows * (unsigned long long)p->width * (unsigned long long)(subset ? p->sample_count : p->spp);
            stats->segments = seg;
            // linear scan: every Scene.Hit tests every sphere; BVH: exact tests counted by the kernel; clusters: spheres whose
            // chunk was scanned (pair pre-filter evaluations), box tests counted separately
            stats->sphere_tests = bvh_tests ? bvh_tests : seg * (unsigned long long)ctx->devs[0].n;
            stats->box_tests = (double)box_tests;
            stats->bounds_violations = (double)violations;
            stats->depth_exhausted = exh;
            stats->kernel_ms = kernel_ms;
            stats->trace_kernel_ms = trace_ms;
            stats->trace_launches = (double)trace_launch_count;
            stats->launches = launches;
            stats->n_devices = G;
            stats->total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count();
        }
    } catch (const std::exception& ex) {
        ctx->rendering = false;
        return fail(ctx, TRAY_E_CUDA, ex.what());
    }
    return TRAY_OK;
}

int tray_read_image(tray_ctx* ctx, uint8_t* rgba_out, size_t stride) {
    if (!ctx) return TRAY_E_INVALID;
    std::lock_guard<std::mutex> lock(ctx->mu);
    if (!ctx->have_image) return fail(ctx, TRAY_E_INVALID, "tray_read_image: nothing rendered");
    if (!rgba_out || stride < (size_t)ctx->width * 4) return fail(ctx, TRAY_E_INVALID, "tray_read_image: bad buffer");
    try { copy_out(ctx, rgba_out, stride); } catch (const std::exception& ex) { return fail(ctx, TRAY_E_CUDA, ex.what()); }
    return TRAY_OK;
}

int tray_upload_frame(tray_ctx* ctx, const uint8_t* rgba, size_t stride, int32_t width, int32_t height) {
    if (!ctx) return TRAY_E_INVALID;
    std::lock_guard<std::mutex> lock(ctx->mu);
    if (!rgba || width <= 0 || height <= 0 || stride < (size_t)width * 4) return fail(ctx, TRAY_E_INVALID, "tray_upload_frame: bad argument");
    try {
        Device& d = ctx->devs[0];
        CK(cudaSetDevice(d.dev));
        const size_t row_bytes = (size_t)width * 4;
        grow(d.rgba, d.rgba_cap, row_bytes * height);
        CK(cudaMemcpy2D(d.rgba, row_bytes, rgba, stride, row_bytes, (size_t)height, cudaMemcpyHostToDevice));
        for (Device& dv : ctx->devs) dv.local_rows.clear();
        for (int y = 0; y < height; y++) d.local_rows.push_back(y);
        ctx->width = width; ctx->height = height; ctx->y0 = 0; ctx->y1 = height;
        ctx->have_image = true; ctx->have_hdr = false; ctx->hdr_is_sums = false;
    } catch (const std::exception& ex) { return fail(ctx, TRAY_E_CUDA, ex.what()); }
    return TRAY_OK;
}

int tray_read_hdr(tray_ctx* ctx, double* hdr_out) {
    if (!ctx) return TRAY_E_INVALID;
    std::lock_guard<std::mutex> lock(ctx->mu);
    if (!ctx->have_hdr || !hdr_out) return fail(ctx, TRAY_E_INVALID, "tray_read_hdr: nothing rendered / null buffer");
    try {
        const size_t row_d = (size_t)ctx->width * 3;
        std::vector<double> tmp;
        for (Device& d : ctx->devs) {
            size_t nrows = d.local_rows.size();
            if (!nrows) continue;
            CK(cudaSetDevice(d.dev));
            tmp.resize(nrows * row_d);
            CK(cudaMemcpy(tmp.data(), d.hdr, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost));
            for (size_t r = 0; r < nrows; r++)
                memcpy(hdr_out + (size_t)d.local_rows[r] * row_d, tmp.data() + r * row_d, row_d * sizeof(double));
        }
    } catch (const std::exception& ex) { return fail(ctx, TRAY_E_CUDA, ex.what()); }
    return TRAY_OK;
}

int tray_cluster_tables(const tray_scene_desc* sc, float* blob_out, size_t cap_floats, int32_t* meta, float* r_out) {
    if (!sc || !meta || sc->n < 0 || sc->n > 65535 || (sc->n > 0 && (!sc->cx || !sc->cy || !sc->cz || !sc->radius))) return TRAY_E_INVALID;
    try {
        ClusterHost H = build_clusters(sc, sc->n);
        meta[0] = (int32_t)H.blob.size(); meta[1] = H.off_box2; meta[2] = H.off_box1; meta[3] = H.off_ids;
        meta[4] = H.real_groups; meta[5] = H.always_groups; meta[6] = (int32_t)H.always_last; meta[7] = H.unfilterable;
        if (r_out) *r_out = H.r;
        if (blob_out) memcpy(blob_out, H.blob.data(), std::min(cap_floats, H.blob.size() * 4) * sizeof(float));
    } catch (const std::exception&) { return TRAY_E_INVALID; }
    return TRAY_OK;
}

int tray_configure(tray_ctx* ctx, int32_t key, int64_t value) {
    if (!ctx) return TRAY_E_INVALID;
    std::lock_guard<std::mutex> lock(ctx->mu);
    if (key == TRAY_CFG_BVH_BUILD && value >= TRAY_BVH_BUILD_AUTO && value <= TRAY_BVH_BUILD_DEVICE) { ctx->bvh_build = (int)value; return TRAY_OK; }
#ifdef TRAY_BOUNDS_CHECK
    if (key == 99) { ctx->debug_fault = (int)value; return TRAY_OK; }  // negative control of the bounds-check build: understate the staged table size
#endif
    if (key == TRAY_CFG_INDISC && value >= 0 && value <= 2) { ctx->indisc_variant = (int)value; return TRAY_OK; }
    if (key == TRAY_CFG_UNITVEC && value >= 0 && value <= 2) { ctx->unitvec_variant = (int)value; return TRAY_OK; }
    return fail(ctx, TRAY_E_INVALID, "tray_configure: unknown key or value");
}

int64_t tray_query(tray_ctx* ctx, int32_t key) {
    if (!ctx) return -1;
    std::lock_guard<std::mutex> lock(ctx->mu);
    if (key == TRAY_CFG_BVH_BUILD) {  // what the builder did for the scene last uploaded: build it now if it is still pending
        try { build_bvh_now(ctx); } catch (const std::exception& ex) { fail(ctx, TRAY_E_CUDA, ex.what()); return -1; }
        return ctx->bvh_built_on_device;
    }
    if (key == TRAY_CFG_INDISC) return ctx->indisc_variant;
    if (key == TRAY_CFG_UNITVEC) return ctx->unitvec_variant;
    return -1;
}

int tray_device_sums(tray_ctx* ctx, double** device_ptr, uint64_t* n_doubles) {
    if (!ctx) return TRAY_E_INVALID;
    std::lock_guard<std::mutex> lock(ctx->mu);
    if (!device_ptr || !n_doubles) return fail(ctx, TRAY_E_INVALID, "tray_device_sums: null argument");
    if (!ctx->have_hdr || !ctx->hdr_is_sums) return fail(ctx, TRAY_E_INVALID, "tray_device_sums: the last render was not a TRAY_SUMS_* render");
    if (ctx->devs.size() != 1) return fail(ctx, TRAY_E_UNSUPPORTED, "tray_device_sums: single-device contexts only (one process per GPU)");
    *device_ptr = ctx->devs[0].hdr;
    *n_doubles = (uint64_t)ctx->devs[0].local_rows.size() * (uint64_t)ctx->width * 3u;
    return TRAY_OK;
}

int tray_resolve_sums(tray_ctx* ctx, uint64_t n_samples, uint8_t* rgba_out, size_t stride) {
    if (!ctx) return TRAY_E_INVALID;
    std::lock_guard<std::mutex> lock(ctx->mu);
    if (!ctx->have_hdr || !ctx->hdr_is_sums) return fail(ctx, TRAY_E_INVALID, "tray_resolve_sums: the last render was not a TRAY_SUMS_* render");
    if (rgba_out && stride < (size_t)ctx->width * 4) return fail(ctx, TRAY_E_INVALID, "tray_resolve_sums: stride < 4*width");
    if (n_samples == 0) n_samples = ctx->sums_samples;
    if (n_samples == 0) return fail(ctx, TRAY_E_INVALID, "tray_resolve_sums: no samples");
    try {
        for (Device& d : ctx->devs) {
            const unsigned long long n_pixels = (unsigned long long)d.local_rows.size() * ctx->width;
            if (!n_pixels) continue;
            CK(cudaSetDevice(d.dev));
            CombineArgs C;
            C.partial[0] = d.hdr; C.n_parts = 1; C.n_pixels = n_pixels;
            C.inv_spp = 1.0 / (double)n_samples;  // colorSumDiv, ray/tracer.go:123
            C.rgba = d.rgba; C.hdr = nullptr; C.srgb_thr = d.srgb_thr;
            combine_kernel<<<(unsigned)((n_pixels + 255) / 256), 256, 0, d.stream>>>(C);
            CK(cudaGetLastError());
        }
        for (Device& d : ctx->devs) { CK(cudaSetDevice(d.dev)); CK(cudaStreamSynchronize(d.stream)); }
        ctx->have_image = true;
        if (rgba_out) copy_out(ctx, rgba_out, stride);
    } catch (const std::exception& ex) { return fail(ctx, TRAY_E_CUDA, ex.what()); }
    return TRAY_OK;
}

int tray_first_hit(tray_ctx* ctx, const tray_camera* cam, int32_t width, int32_t height, int32_t precision,
                   int32_t* id, double* t, double* normal, uint8_t* front) {
    if (!ctx) return TRAY_E_INVALID;
    std::lock_guard<std::mutex> lock(ctx->mu);
    if (!cam || !id || !t || !normal || !front || width <= 0 || height <= 0) return fail(ctx, TRAY_E_INVALID, "tray_first_hit: bad argument");
    if (!ctx->have_scene) return fail(ctx, TRAY_E_NO_SCENE, "tray_first_hit: no scene uploaded");
    try {
        Device& d = ctx->devs[0];
        CK(cudaSetDevice(d.dev));
        size_t n = (size_t)width * height;
        DevTmp t_id(n * 4), t_t(n * 8), t_n(n * 24), t_f(n);  // freed on every exit path
        int* d_id = t_id.as<int>(); double* d_t = t_t.as<double>(); double* d_n = t_n.as<double>(); unsigned char* d_f = t_f.as<unsigned char>();
        DevCamera dc = dev_camera(cam);
        unsigned blocks = (unsigned)((n + 127) / 128);
        if (precision == TRAY_FP64_FMA) first_hit_kernel<double, true><<<blocks, 128, 0, d.stream>>>(dc, dev_scene<double>(ctx, d), width, height, d_id, d_t, d_n, d_f);
        else if (precision == TRAY_FP64_STRICT || precision == TRAY_FP64_STRICT_BRUTE) first_hit_kernel<double, false><<<blocks, 128, 0, d.stream>>>(dc, dev_scene<double>(ctx, d), width, height, d_id, d_t, d_n, d_f);
        else first_hit_kernel<float, true><<<blocks, 128, 0, d.stream>>>(dc, dev_scene<float>(ctx, d), width, height, d_id, d_t, d_n, d_f);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(d.stream));
        CK(cudaMemcpy(id, d_id, n * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(t, d_t, n * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(normal, d_n, n * 24, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(front, d_f, n, cudaMemcpyDeviceToHost));
    } catch (const std::exception& ex) { return fail(ctx, TRAY_E_CUDA, ex.what()); }
    return TRAY_OK;
}

int tray_rng_dump(tray_ctx* ctx, int32_t kind, uint64_t idx, uint64_t seed, double radius, int32_t n, double* out) {
    if (!ctx) return TRAY_E_INVALID;
    std::lock_guard<std::mutex> lock(ctx->mu);
    if (!out || n <= 0 || kind < 0 || kind > 4) return fail(ctx, TRAY_E_INVALID, "tray_rng_dump: bad argument");
    try {
        Device& d = ctx->devs[0];
        CK(cudaSetDevice(d.dev));
        int per = kind == 3 ? 3 : (kind == 4 ? 2 : 1);
        DevTmp tmp(sizeof(double) * n * per);
        double* dout = tmp.as<double>();
        rng_dump_kernel<<<1, 128, 0, d.stream>>>(kind, idx, seed, radius, n, dout, ctx->indisc_variant, ctx->unitvec_variant);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(d.stream));
        CK(cudaMemcpy(out, dout, sizeof(double) * n * per, cudaMemcpyDeviceToHost));
    } catch (const std::exception& ex) { return fail(ctx, TRAY_E_CUDA, ex.what()); }
    return TRAY_OK;
}

int tray_arith_probe(tray_ctx* ctx, int32_t kind, const double* a, const double* b, int32_t n, double* out) {
    if (!ctx) return TRAY_E_INVALID;
    std::lock_guard<std::mutex> lock(ctx->mu);
    if (!a || !b || !out || n <= 0 || kind < 0 || kind > 3) return fail(ctx, TRAY_E_INVALID, "tray_arith_probe: bad argument");
    try {
        Device& d = ctx->devs[0];
        CK(cudaSetDevice(d.dev));
        const int per = kind == 0 ? 3 : 1;
        DevTmp t_a(sizeof(double) * n), t_b(sizeof(double) * n), t_o(sizeof(double) * n * per);
        CK(cudaMemcpy(t_a.as<double>(), a, sizeof(double) * n, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(t_b.as<double>(), b, sizeof(double) * n, cudaMemcpyHostToDevice));
        arith_probe_kernel<<<(n + 255) / 256, 256, 0, d.stream>>>(kind, t_a.as<double>(), t_b.as<double>(), n, t_o.as<double>());
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(d.stream));
        CK(cudaMemcpy(out, t_o.as<double>(), sizeof(double) * n * per, cudaMemcpyDeviceToHost));
    } catch (const std::exception& ex) { return fail(ctx, TRAY_E_CUDA, ex.what()); }
    return TRAY_OK;
}

int tray_linear_to_srgb(tray_ctx* ctx, const double* x, int32_t n, uint8_t* out) {
    if (!ctx) return TRAY_E_INVALID;
    std::lock_guard<std::mutex> lock(ctx->mu);
    if (!x || !out || n <= 0) return fail(ctx, TRAY_E_INVALID, "tray_linear_to_srgb: bad argument");
    try {
        Device& d = ctx->devs[0];
        CK(cudaSetDevice(d.dev));
        DevTmp t_x(sizeof(double) * n), t_o(n);
        double* dx = t_x.as<double>(); unsigned char* dout = t_o.as<unsigned char>();
        CK(cudaMemcpy(dx, x, sizeof(double) * n, cudaMemcpyHostToDevice));
        srgb_kernel<<<(n + 255) / 256, 256, 0, d.stream>>>(dx, n, d.srgb_thr, dout);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(d.stream));
        CK(cudaMemcpy(out, dout, n, cudaMemcpyDeviceToHost));
    } catch (const std::exception& ex) { return fail(ctx, TRAY_E_CUDA, ex.what()); }
    return TRAY_OK;
}

int tray_present(tray_ctx* ctx, int32_t cols, int32_t rows2, uint8_t* rgba_small_out, uint8_t* ansi_out, size_t ansi_cap,
                 size_t* ansi_len, double* device_ms) {
    if (!ctx) return TRAY_E_INVALID;
    std::lock_guard<std::mutex> lock(ctx->mu);
    if (!ctx->have_image) return fail(ctx, TRAY_E_INVALID, "tray_present: nothing rendered");
    if (cols <= 0 || rows2 <= 0 || (rows2 & 1)) return fail(ctx, TRAY_E_INVALID, "tray_present: cols > 0 and an even rows2 > 0 required");
    Device& d = ctx->devs[0];
    if (ctx->y0 != 0 || ctx->y1 != ctx->height || (int)d.local_rows.size() != ctx->height)
        return fail(ctx, TRAY_E_UNSUPPORTED, "tray_present: needs a complete frame resident on one device (single-device context or sample split)");
    const size_t need = (size_t)(rows2 / 2) * ((size_t)cols * kAnsiCell + kAnsiEol);
    if (ansi_len) *ansi_len = need;
    if (ansi_out && ansi_cap < need) return fail(ctx, TRAY_E_INVALID, "tray_present: ansi_out too small");
    try {
        CK(cudaSetDevice(d.dev));
        const int sw = ctx->width, sh = ctx->height;
        grow(d.pr_tmp, d.pr_tmp_cap, (size_t)cols * sh);
        grow(d.pr_small, d.pr_small_cap, (size_t)cols * rows2);
        grow(d.pr_ansi, d.pr_ansi_cap, need);
        double4* tmp = d.pr_tmp; uchar4* small = d.pr_small; unsigned char* ansi = d.pr_ansi;
        cudaEvent_t a = next_event(d), b = next_event(d);
        CK(cudaEventRecord(a, d.stream));
        if (sw < cols) {  // supersample < 1: the frame is smaller than the terminal -> draw.NearestNeighbor (main.go:124-125)
            scale_nn_kernel<<<dim3((cols + 127) / 128, rows2), 128, 0, d.stream>>>(reinterpret_cast<const uchar4*>(d.rgba), sw, sh, cols, rows2, small);
        } else {          // draw.BiLinear (main.go:126-127); identity when the sizes agree (supersample == 1: no scaling at all)
            scale_x_kernel<<<dim3((cols + 127) / 128, sh), 128, 0, d.stream>>>(reinterpret_cast<const uchar4*>(d.rgba), sw, sh, cols, tmp);
            scale_y_kernel<<<dim3((cols + 127) / 128, rows2), 128, 0, d.stream>>>(tmp, cols, sh, rows2, small);
        }
        ansi_kernel<<<dim3((cols + 1 + 127) / 128, rows2 / 2), 128, 0, d.stream>>>(small, cols, rows2 / 2, ansi);
        CK(cudaGetLastError());
        CK(cudaEventRecord(b, d.stream));
        if (ansi_out) CK(cudaMemcpyAsync(ansi_out, ansi, need, cudaMemcpyDeviceToHost, d.stream));
        if (rgba_small_out) CK(cudaMemcpyAsync(rgba_small_out, small, sizeof(uchar4) * (size_t)cols * rows2, cudaMemcpyDeviceToHost, d.stream));
        CK(cudaStreamSynchronize(d.stream));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (device_ms) *device_ms = ms;
    } catch (const std::exception& ex) { return fail(ctx, TRAY_E_CUDA, ex.what()); }
    return TRAY_OK;
}

static uint32_t host_crc32(const uint8_t* p, size_t n) {
    uint32_t c = 0xFFFFFFFFu;
    for (size_t i = 0; i < n; i++) {
        c ^= p[i];
        for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
    }
    return ~c;
}

size_t tray_png_bound(int32_t width, int32_t height) {
    if (width <= 0 || height <= 0) return 0;
    const size_t raw = (size_t)height * (1 + 3 * (size_t)width);
    return 2 * raw + 1024 * 1024;  // 15-bit codes at worst (< 2x), block headers, file framing
}

int tray_encode_png(tray_ctx* ctx, uint8_t* png_out, size_t cap, size_t* png_len, double* device_ms) {
    if (!ctx) return TRAY_E_INVALID;
    std::lock_guard<std::mutex> lock(ctx->mu);
    if (!ctx->have_image) return fail(ctx, TRAY_E_INVALID, "tray_encode_png: nothing rendered");
    Device& d = ctx->devs[0];
    size_t have_rows = 0;
    for (Device& dv : ctx->devs) have_rows += dv.local_rows.size();
    if (ctx->y0 != 0 || ctx->y1 != ctx->height || (int)have_rows != ctx->height)
        return fail(ctx, TRAY_E_UNSUPPORTED, "tray_encode_png: needs a complete frame in this context (all rows, no external shards)");
    if (!png_len) return fail(ctx, TRAY_E_INVALID, "tray_encode_png: png_len is NULL");
    try {
        CK(cudaSetDevice(d.dev));
        // tile mode on several devices: gather the row bands onto device 0 (peer copies over NVLink), still no host pixels
        const uchar4* frame = reinterpret_cast<const uchar4*>(d.rgba);
        unsigned char* gathered = nullptr;
        std::unique_ptr<DevTmp> gathered_mem;  // freed on every exit path
        if ((int)d.local_rows.size() != ctx->height) {
            const size_t row_bytes = (size_t)ctx->width * 4;
            gathered_mem.reset(new DevTmp(row_bytes * ctx->height));
            gathered = gathered_mem->as<unsigned char>();
            for (Device& dv : ctx->devs) {
                size_t r = 0, nrows = dv.local_rows.size();
                while (r < nrows) {
                    size_t e = r + 1;
                    while (e < nrows && dv.local_rows[e] == dv.local_rows[e - 1] + 1) e++;
                    CK(cudaMemcpyPeerAsync(gathered + (size_t)dv.local_rows[r] * row_bytes, d.dev, dv.rgba + r * row_bytes, dv.dev, (e - r) * row_bytes, d.stream));
                    r = e;
                }
            }
            frame = reinterpret_cast<const uchar4*>(gathered);
        }
        PngPlan P;
        P.width = ctx->width; P.height = ctx->height; P.row_len = 1 + 3 * ctx->width;
        // deflate blocks: >= 32 KB of scanlines each, about one block per SM when the image is large enough
        int rpb = std::max((32768 + P.row_len - 1) / P.row_len, P.height / std::max(1, d.num_sms));
        P.rows_per_block = std::max(1, rpb);
        P.n_blocks = (P.height + P.rows_per_block - 1) / P.rows_per_block;
        const size_t raw = (size_t)P.height * P.row_len;
        const size_t out_cap = (tray_png_bound(P.width, P.height) + 15) / 16 * 16;
        const size_t cap_pieces = out_cap / kCrcChunk + 2;
        grow(d.png_filt, d.png_filt_cap, raw); grow(d.png_out, d.png_out_cap, out_cap);
        grow(d.png_hist, d.png_hist_cap, (size_t)256 * P.n_blocks); grow(d.png_piece, d.png_piece_cap, cap_pieces);
        grow(d.png_adler, d.png_adler_cap, (size_t)2 * P.height);
        grow(d.png_blocks, d.png_blocks_cap, (size_t)P.n_blocks); grow(d.png_tot, d.png_tot_cap, (size_t)1);
        unsigned char *filt = d.png_filt, *out = d.png_out;
        unsigned *hist = d.png_hist, *piece = d.png_piece;
        unsigned long long* row_adler = d.png_adler;
        PngBlock* blocks = d.png_blocks;
        PngTotals* tot = d.png_tot;
        // the fixed 43 bytes in front of the deflate stream: signature, IHDR chunk, IDAT length (patched on the device), "IDAT", zlib header
        uint8_t pre[43] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n', 0, 0, 0, 13, 'I', 'H', 'D', 'R'};
        auto be32 = [](uint8_t* q, uint32_t v) { q[0] = v >> 24; q[1] = v >> 16; q[2] = v >> 8; q[3] = v; };
        be32(pre + 16, (uint32_t)P.width); be32(pre + 20, (uint32_t)P.height);
        pre[24] = 8; pre[25] = 2; pre[26] = 0; pre[27] = 0; pre[28] = 0;  // 8 bit, truecolour, deflate, adaptive filtering, no interlace
        be32(pre + 29, host_crc32(pre + 12, 17));
        be32(pre + 33, 0);
        memcpy(pre + 37, "IDAT", 4);
        pre[41] = 0x78; pre[42] = 0x01;
        cudaEvent_t a = next_event(d), b = next_event(d);
        CK(cudaEventRecord(a, d.stream));
        CK(cudaMemsetAsync(out, 0, out_cap, d.stream));
        CK(cudaMemsetAsync(hist, 0, sizeof(unsigned) * 256 * P.n_blocks, d.stream));
        CK(cudaMemcpyAsync(out, pre, sizeof pre, cudaMemcpyHostToDevice, d.stream));
        png_filter_kernel<<<P.height, 256, 0, d.stream>>>(frame, P, filt, hist, row_adler);
        png_huffman_kernel<<<P.n_blocks, 320, 0, d.stream>>>(hist, P, blocks);
        png_layout_kernel<<<1, 32, 0, d.stream>>>(P, blocks, row_adler, tot);
        png_pack_kernel<<<P.n_blocks, kPackThreads, 0, d.stream>>>(filt, P, blocks, reinterpret_cast<unsigned*>(out));
        png_finish_kernel<<<1, 256, 0, d.stream>>>(out, tot, piece, 0);
        png_crc_kernel<<<(unsigned)((cap_pieces + 255) / 256), 256, 0, d.stream>>>(out, tot, piece, cap_pieces);
        png_finish_kernel<<<1, 256, 0, d.stream>>>(out, tot, piece, 1);
        CK(cudaGetLastError());
        CK(cudaEventRecord(b, d.stream));
        PngTotals ht;
        CK(cudaMemcpyAsync(&ht, tot, sizeof ht, cudaMemcpyDeviceToHost, d.stream));
        CK(cudaStreamSynchronize(d.stream));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (device_ms) *device_ms = ms;
        *png_len = (size_t)ht.file_bytes;
        int rc = TRAY_OK;
        if (ht.file_bytes > out_cap) rc = fail(ctx, TRAY_E_CUDA, "tray_encode_png: internal bound exceeded");
        else if (png_out) {
            if (cap < ht.file_bytes) rc = fail(ctx, TRAY_E_INVALID, "tray_encode_png: png_out too small (png_len holds the size needed)");
            else CK(cudaMemcpy(png_out, out, (size_t)ht.file_bytes, cudaMemcpyDeviceToHost));
        }
        return rc;
    } catch (const std::exception& ex) { return fail(ctx, TRAY_E_CUDA, ex.what()); }
}

uint64_t tray_progress(tray_ctx* ctx) {
    if (!ctx) return 0;
    if (!ctx->rendering.load()) return ctx->progress_base.load();
    // render in flight (another thread holds ctx->mu): peek at the device counters on each device's side stream
    std::lock_guard<std::mutex> lock(ctx->peek_mu);
    uint64_t samples = 0;
    for (Device& d : ctx->devs) {
        if (cudaSetDevice(d.dev) != cudaSuccess) continue;
        if (cudaMemcpyAsync(d.peek_host, d.stats + 2, sizeof(unsigned long long), cudaMemcpyDeviceToHost, d.peek_stream) != cudaSuccess) continue;
        if (cudaStreamSynchronize(d.peek_stream) != cudaSuccess) continue;
        samples += *d.peek_host;
    }
    int spp = ctx->progress_spp.load();
    return samples / (uint64_t)(spp > 0 ? spp : 1);
}

int tray_measure_peak(tray_ctx* ctx, int32_t kind, double* tflops, double* ms_out) {
    if (!ctx) return TRAY_E_INVALID;
    std::lock_guard<std::mutex> lock(ctx->mu);
    if (!tflops || kind < 0 || kind > 5) return fail(ctx, TRAY_E_INVALID, "tray_measure_peak: bad argument");
    if (kind == 3 || kind == 4) {  // hot-loop-only probe on the uploaded scene (3: fused, 4: strict); TFLOP/s at 18 flops per test
        if (!ctx->have_scene) return fail(ctx, TRAY_E_NO_SCENE, "tray_measure_peak: loop probe needs a scene");
        try {
            Device& d = ctx->devs[0];
            CK(cudaSetDevice(d.dev));
            double* sink;
            CK(cudaMalloc(&sink, 8));
            DevScene<double> S = dev_scene<double>(ctx, d);
            size_t smem = (size_t)S.n_pad * sizeof(double4);
            if (smem > kSmemBudget) throw std::runtime_error("tray_measure_peak: scene too large for the loop probe");
            const int iters = 200;
            int bps = 0;
            cudaEvent_t a, b;
            CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
            float best = 1e30f;
            int blocks = 0;
            for (int rep = 0; rep < 3; rep++) {
                if (kind == 3) {
                    auto k = hotloop_probe_kernel<double, true, kTPB, kMinBlocks>;
                    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k, kTPB, smem));
                    blocks = d.num_sms * std::max(1, bps);
                    CK(cudaEventRecord(a, d.stream));
                    k<<<blocks, kTPB, smem, d.stream>>>(S, iters, sink);
                } else {
                    auto k = hotloop_probe_kernel<double, false, kTPB, kMinBlocks>;
                    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k, kTPB, smem));
                    blocks = d.num_sms * std::max(1, bps);
                    CK(cudaEventRecord(a, d.stream));
                    k<<<blocks, kTPB, smem, d.stream>>>(S, iters, sink);
                }
                CK(cudaGetLastError());
                CK(cudaEventRecord(b, d.stream));
                CK(cudaStreamSynchronize(d.stream));
                float ms;
                CK(cudaEventElapsedTime(&ms, a, b));
                if (rep > 0) best = std::min(best, ms);
            }
            *tflops = 18.0 * (double)S.n * iters * (double)blocks * kTPB / (best * 1e-3) / 1e12;
            if (ms_out) *ms_out = best;
            cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(sink);
        } catch (const std::exception& ex) { return fail(ctx, TRAY_E_CUDA, ex.what()); }
        return TRAY_OK;
    }
    try {
        Device& d = ctx->devs[0];
        CK(cudaSetDevice(d.dev));
        double* sink;
        CK(cudaMalloc(&sink, 8));
        const int iters = 1 << 15, blocks = d.num_sms * 8, tpb = 256;
        cudaEvent_t a, b;
        CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
        float best = 1e30f;
        for (int rep = 0; rep < 4; rep++) {
            CK(cudaEventRecord(a, d.stream));
            if (kind == 0) peak_kernel<0><<<blocks, tpb, 0, d.stream>>>(sink, iters, 1e-9);
            else if (kind == 1) peak_kernel<1><<<blocks, tpb, 0, d.stream>>>(sink, iters, 1e-9);
            else if (kind == 5) peak_kernel<5><<<blocks, tpb, 0, d.stream>>>(sink, iters, 1e-9);
            else peak_kernel<2><<<blocks, tpb, 0, d.stream>>>(sink, iters, 1e-9);
            CK(cudaGetLastError());
            CK(cudaEventRecord(b, d.stream));
            CK(cudaStreamSynchronize(d.stream));
            float ms;
            CK(cudaEventElapsedTime(&ms, a, b));
            if (rep > 0) best = std::min(best, ms);
        }
        // ops: every chain step is 2 flops (fma, or mul+add)
        double flops = 2.0 * 8 * (double)iters * (double)blocks * tpb * (kind == 5 ? 2.0 : 1.0);
        *tflops = flops / (best * 1e-3) / 1e12;
        if (ms_out) *ms_out = best;
        cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(sink);
    } catch (const std::exception& ex) { return fail(ctx, TRAY_E_CUDA, ex.what()); }
    return TRAY_OK;
}

}  // extern "C"
