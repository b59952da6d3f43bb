/*
 * tray_oracle.c -- CPU ORACLE for the fortio/tray path-tracing hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE. Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it. The product
 * (libtraycuda.so) never links, loads or calls anything in oracle/.
 *
 * It is a plain-C restatement (no code copied) of the reference's Go algorithm,
 * operation by operation, in the reference's evaluation order, compiled with
 *     gcc -O2 -ffp-contract=off -fno-fast-math
 * so that x86-64 SSE2 doubles reproduce Go/amd64 float64 semantics (Go on amd64 never
 * fuses x*y+z). Each function cites the reference file:line it follows
 * (paths relative to /root/reference).
 *
 * The Go reference itself cannot be built here (no go/gccgo toolchain), so there is no
 * oracle/_ref. Third-party arithmetic that is NOT in the reference tree is restated from
 * the published algorithms:
 *   - Go math/rand/v2 PCG-DXSM, Float64, NormFloat64 (via fortio.org/rand v1.1.0, go.mod:9)
 *   - fortio.org/rand wrappers New/NewIdx/Float64Range/Vec3/UnitVector/InDisc
 *   - fortio.org/terminal v0.63.4 tcolor.LinearToSrgb (go.mod:10)
 *   - Go math.Log / math.Exp / math.Pow / math.Tan (pure-Go FreeBSD-msun / Cephes forms)
 * PIN STATUS (see DESIGN.md "Oracle pins"): PCG is pinned by Go's own known-answer vector;
 * seeding + Float64 by the "486 objects at seed 7" comment (benchmark/benchmark.go:42);
 * LinearToSrgb by ray/vec3_test.go:264-289; the ziggurat table heads by Go's literals.
 * UnitVector / InDisc bodies are PARITY UNPINNED at value level (no golden vector in the
 * reference pins them; the example.png experiment in tools/pin_example_png.py is the
 * strongest available evidence and its outcome is recorded in DESIGN.md).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "zig_tables.h"

#define EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------ */
/* Go math restatements (deterministic, only + - * / and bit ops)       */
/* ------------------------------------------------------------------ */

/* Go math.Frexp for finite normal x>0: frac in [0.5,1), x = frac * 2^exp. */
static double go_frexp(double x, int *e) { return frexp(x, e); /* exact by definition */ }

/* Go math.Log, pure-Go form (FreeBSD e_log.c): src/math/log.go. */
static double go_log(double x) {
    const double Ln2Hi = 6.93147180369123816490e-01, Ln2Lo = 1.90821492927058770002e-10;
    const double L1 = 6.666666666666735130e-01, L2 = 3.999999999940941908e-01,
                 L3 = 2.857142874366239149e-01, L4 = 2.222219843214978396e-01,
                 L5 = 1.818357216161805012e-01, L6 = 1.531383769920937332e-01,
                 L7 = 1.479819860511658591e-01;
    if (isnan(x) || (isinf(x) && x > 0)) return x;
    if (x < 0) return NAN;
    if (x == 0) return -INFINITY;
    int ki;
    double f1 = go_frexp(x, &ki);
    if (f1 < 0.70710678118654752440 /* Sqrt2/2 */) { f1 *= 2; ki--; }
    double f = f1 - 1;
    double k = (double)ki;
    double s = f / (2 + f);
    double s2 = s * s;
    double s4 = s2 * s2;
    double t1 = s2 * (L1 + s4 * (L3 + s4 * (L5 + s4 * L7)));
    double t2 = s4 * (L2 + s4 * (L4 + s4 * L6));
    double R = t1 + t2;
    double hfsq = 0.5 * f * f;
    return k * Ln2Hi - ((hfsq - (s * (hfsq + R) + k * Ln2Lo)) - f);
}

/* Go math.Exp, pure-Go form (FreeBSD e_exp.c): src/math/exp.go. */
static double go_exp(double x) {
    const double Ln2Hi = 6.93147180369123816490e-01, Ln2Lo = 1.90821492927058770002e-10,
                 Log2e = 1.44269504088896338700e+00;
    const double Overflow = 7.09782712893383973096e+02, Underflow = -7.45133219101941108420e+02,
                 NearZero = 1.0 / (1 << 28);
    const double P1 = 1.66666666666666657415e-01, P2 = -2.77777777770155933842e-03,
                 P3 = 6.61375632143793436117e-05, P4 = -1.65339022054652515390e-06,
                 P5 = 4.13813679705723846039e-08;
    if (isnan(x) || (isinf(x) && x > 0)) return x;
    if (isinf(x)) return 0;
    if (x > Overflow) return INFINITY;
    if (x < Underflow) return 0;
    if (-NearZero < x && x < NearZero) return 1 + x;
    int k = 0;
    if (x < 0) k = (int)(Log2e * x - 0.5);
    else if (x > 0) k = (int)(Log2e * x + 0.5);
    double hi = x - (double)k * Ln2Hi;
    double lo = (double)k * Ln2Lo;
    double r = hi - lo;
    double t = r * r;
    double c = r - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
    double y = 1 - ((lo - (r * c) / (2 - c)) - hi);
    return ldexp(y, k);
}

/* Go math.Pow(x, 5) for finite x >= 0: frexp + square-and-multiply (src/math/pow.go);
 * power-of-two renormalisations are exact, so the roundings are those of x*((x*x)*(x*x)). */
static double go_pow5(double x) {
    double x2 = x * x;
    return x * (x2 * x2);
}

/* Go math.Pow(x, y) for 0 < x < 1, 0 < y < 0.5 non-integer: yi = 0, so Pow = Exp(y*Log(x)). */
static double go_pow_frac(double x, double y) { return go_exp(y * go_log(x)); }

/* Go math.Tan, pure-Go Cephes form (src/math/tan.go), |x| < 2^29 path. */
static double go_tan(double x) {
    const double PI4A = 7.85398125648498535156e-1, PI4B = 3.77489470793079817668e-8,
                 PI4C = 2.69515142907905952645e-15;
    const double P0 = -1.30936939181383777646e4, P1 = 1.15351664838587416140e6,
                 P2 = -1.79565251976484877988e7;
    const double Q1 = 1.36812963470692954678e4, Q2 = -1.32089234440210967447e6,
                 Q3 = 2.50083801823357915839e7, Q4 = -5.38695755929454629881e7;
    int sign = 0;
    if (x == 0 || isnan(x)) return x;
    if (isinf(x)) return NAN;
    if (x < 0) { x = -x; sign = 1; }
    uint64_t j = (uint64_t)(x * (4 / M_PI));
    double y = (double)j;
    if (j & 1) { j++; y++; }
    double z = ((x - y * PI4A) - y * PI4B) - y * PI4C;
    double zz = z * z;
    if (zz > 1e-14)
        y = z + z * (zz * (((P0 * zz) + P1) * zz + P2) / ((((zz + Q1) * zz + Q2) * zz + Q3) * zz + Q4));
    else
        y = z;
    if (j & 2) y = -1 / y;
    if (sign) y = -y;
    return y;
}

/* Go math.Sin / math.Cos, pure-Go Cephes forms (src/math/sin.go; no assembly on amd64 / arm64), arguments below
 * the Payne-Hanek threshold (the variants below call them with angles in [0, 2*Pi)). Only + - * and exact
 * conversions, so the device restatement (tray_device.cuh) agrees bit for bit. */
static const double GO_SIN[6] = {1.58962301576546568060e-10, -2.50507477628578072866e-8, 2.75573136213857245213e-6,
                                 -1.98412698295895385996e-4, 8.33333333332211858878e-3, -1.66666666666666307295e-1};
static const double GO_COS[6] = {-1.13585365213876817300e-11, 2.08757008419747316778e-9, -2.75573141792967388112e-7,
                                 2.48015872888517045348e-5, -1.38888888888730564116e-3, 4.16666666666665929218e-2};
static double go_sincos_reduce(double x, uint64_t *jo) {
    const double PI4A = 7.85398125648498535156e-1, PI4B = 3.77489470793079817668e-8,
                 PI4C = 2.69515142907905952645e-15;
    uint64_t j = (uint64_t)(x * (4 / M_PI));
    double y = (double)j;
    if (j & 1) { j++; y++; }
    *jo = j & 7;
    return ((x - y * PI4A) - y * PI4B) - y * PI4C;
}
static double go_sin_poly(double z, double zz) {
    return z + z * zz * ((((((GO_SIN[0] * zz) + GO_SIN[1]) * zz + GO_SIN[2]) * zz + GO_SIN[3]) * zz + GO_SIN[4]) * zz + GO_SIN[5]);
}
static double go_cos_poly(double zz) {
    return 1.0 - 0.5 * zz + zz * zz * ((((((GO_COS[0] * zz) + GO_COS[1]) * zz + GO_COS[2]) * zz + GO_COS[3]) * zz + GO_COS[4]) * zz + GO_COS[5]);
}
static double go_sin(double x) {
    int sign = 0;
    if (x == 0 || isnan(x)) return x;
    if (isinf(x)) return NAN;
    if (x < 0) { x = -x; sign = 1; }
    uint64_t j;
    double z = go_sincos_reduce(x, &j);
    if (j > 3) { sign = !sign; j -= 4; }
    double zz = z * z;
    double y = (j == 1 || j == 2) ? go_cos_poly(zz) : go_sin_poly(z, zz);
    return sign ? -y : y;
}
static double go_cos(double x) {
    int sign = 0;
    if (isnan(x) || isinf(x)) return NAN;
    x = fabs(x);
    uint64_t j;
    double z = go_sincos_reduce(x, &j);
    if (j > 3) { j -= 4; sign = !sign; }
    if (j > 1) sign = !sign;
    double zz = z * z;
    double y = (j == 1 || j == 2) ? go_sin_poly(z, zz) : go_cos_poly(zz);
    return sign ? -y : y;
}

/* ------------------------------------------------------------------ */
/* RNG: Go math/rand/v2 PCG-DXSM + fortio.org/rand wrappers              */
/* ------------------------------------------------------------------ */

typedef struct { uint64_t hi, lo; uint64_t draws; } rng_t;

static int g_indisc_variant = 0;   /* 0 rejection, 1 polar(angle,r), 2 polar(r,angle) */
static int g_unitvec_variant = 0;  /* 0 normals, 1 cube rejection, 2 angle */

EXPORT void oracle_set_variants(int indisc, int unitvec) { g_indisc_variant = indisc; g_unitvec_variant = unitvec; }

/* rand.NewIdx(idx, seed) == rand.NewPCG(uint64(idx), seed); rand.New(seed) == NewIdx(0, seed)
 * (call sites ray/tracer.go:121, benchmark/benchmark.go:63). */
static int g_seed_variant = 0; /* experiment hook (tools/pin_example_png.py); 0 = default */
static int g_xlimit = 0;       /* experiment hook: render only x < g_xlimit of each row */
static rng_t rng_new_idx(uint64_t idx, uint64_t seed) {
    rng_t r = {idx, seed, 0};
    if (g_seed_variant == 1) { r.hi = seed; r.lo = idx; }
    else if (g_seed_variant == 2) { r.hi = 0; r.lo = seed + idx; }
    else if (g_seed_variant == 3) { r.hi = seed + idx; r.lo = 0; }
    else if (g_seed_variant == 4) { r.hi = seed; r.lo = seed + idx; }
    return r;
}
EXPORT void oracle_set_experiment(int seed_variant, int xlimit) { g_seed_variant = seed_variant; g_xlimit = xlimit; }

/* math/rand/v2 (*PCG).Uint64: 128-bit LCG step then DXSM output on the NEW state. */
static uint64_t rng_u64(rng_t *r) {
    const unsigned __int128 MUL = ((unsigned __int128)2549297995355413924ULL << 64) | 4865540595714422341ULL;
    const unsigned __int128 INC = ((unsigned __int128)6364136223846793005ULL << 64) | 1442695040888963407ULL;
    unsigned __int128 s = ((unsigned __int128)r->hi << 64) | r->lo;
    s = s * MUL + INC;
    r->hi = (uint64_t)(s >> 64);
    r->lo = (uint64_t)s;
    r->draws++;
    uint64_t hi = r->hi, lo = r->lo;
    hi ^= hi >> 32;
    hi *= 0xda942042e4dd58b5ULL;
    hi ^= hi >> 48;
    hi *= (lo | 1);
    return hi;
}

/* math/rand/v2 (*Rand).Float64: float64(Uint64()<<11>>11) / (1<<53). */
static double rng_f64(rng_t *r) { return (double)(rng_u64(r) << 11 >> 11) / 9007199254740992.0; }

static const uint32_t *zig_kn = ZIG_KN_INIT;
static const float *zig_wn = ZIG_WN_INIT;
static const float *zig_fn = ZIG_FN_INIT;

/* math/rand/v2 (*Rand).NormFloat64 (ziggurat, 128 strips), SURVEY App. A.3. */
static double rng_norm(rng_t *r) {
    for (;;) {
        uint64_t u = rng_u64(r);
        int32_t j = (int32_t)(uint32_t)u;
        uint32_t i = (uint32_t)(u >> 32) & 0x7F;
        double x = (double)j * (double)zig_wn[i];
        uint32_t aj = j < 0 ? (uint32_t)(-(int64_t)j) : (uint32_t)j;
        if (aj < zig_kn[i]) return x;
        if (i == 0) {
            for (;;) {
                x = -go_log(rng_f64(r)) * ZIG_INV_RN;
                double y = -go_log(rng_f64(r));
                if (y + y >= x * x) break;
            }
            if (j > 0) return ZIG_RN + x;
            return -ZIG_RN - x;
        }
        float lhs = zig_fn[i] + (float)rng_f64(r) * (zig_fn[i - 1] - zig_fn[i]);
        if (lhs < (float)go_exp(-.5 * x * x)) return x;
    }
}

/* fortio.org/rand Rand.UnitVector (called from ray/rand.go:31). "Norm method"
 * (ray/vec3_test.go:513,550): three normals, normalised. */
static void rng_unit_vector(rng_t *r, double *ox, double *oy, double *oz) {
    if (g_unitvec_variant == 1) { /* ray/rand.go:50-58 (legacy rejection) */
        for (;;) {
            double x = -1 + (1 - -1) * rng_f64(r), y = -1 + 2 * rng_f64(r), z = -1 + 2 * rng_f64(r);
            double l2 = x * x + y * y + z * z;
            if (l2 > 1e-48 && l2 <= 1) { double l = sqrt(l2); *ox = x / l; *oy = y / l; *oz = z / l; return; }
        }
    }
    if (g_unitvec_variant == 2) { /* ray/rand.go:62-69 (legacy angle) */
        double angle = rng_f64(r) * 2 * M_PI;
        double z = rng_f64(r) * 2 - 1;
        double rad = sqrt(1 - z * z);
        *ox = rad * go_cos(angle); *oy = rad * go_sin(angle); *oz = z;
        return;
    }
    for (;;) {
        double x = rng_norm(r), y = rng_norm(r), z = rng_norm(r);
        double rad = sqrt(x * x + y * y + z * z);
        if (rad > 1e-24) { *ox = x / rad; *oy = y / rad; *oz = z / rad; return; }
    }
}

/* fortio.org/rand Rand.InDisc(radius) (call sites ray/tracer.go:138, ray/camera.go:128).
 * Body not in the reference tree; variant 0 (rejection in the square, 2 draws per try)
 * is the default, see DESIGN.md. */
static void rng_in_disc(rng_t *r, double radius, double *ox, double *oy) {
    if (g_indisc_variant == 1 || g_indisc_variant == 2) {
        double u1 = rng_f64(r), u2 = rng_f64(r);
        double ang = (g_indisc_variant == 1 ? u1 : u2) * 2 * M_PI;
        double rr = radius * sqrt(g_indisc_variant == 1 ? u2 : u1);
        *ox = rr * go_cos(ang); *oy = rr * go_sin(ang);
        return;
    }
    for (;;) {
        double x = 2 * rng_f64(r) - 1;
        double y = 2 * rng_f64(r) - 1;
        if (x * x + y * y <= 1) { *ox = radius * x; *oy = radius * y; return; }
    }
}

/* ------------------------------------------------------------------ */
/* Vec3 (ray/vec3.go:25-145)                                             */
/* ------------------------------------------------------------------ */
typedef struct { double x, y, z; } v3;
static inline v3 V(double x, double y, double z) { v3 r = {x, y, z}; return r; }
static inline v3 v_add(v3 u, v3 v) { return V(v.x + u.x, v.y + u.y, v.z + u.z); }       /* :25 */
static inline v3 v_sub(v3 u, v3 v) { return V(u.x - v.x, u.y - v.y, u.z - v.z); }       /* :30 */
static inline double v_dot(v3 u, v3 v) { return u.x * v.x + u.y * v.y + u.z * v.z; }    /* :58 */
static inline v3 v_cross(v3 u, v3 v) {                                                   /* :71 */
    return V(u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x);
}
static inline v3 v_smul(v3 v, double t) { return V(v.x * t, v.y * t, v.z * t); }        /* :92 */
static inline v3 v_mul(v3 u, v3 v) { return V(u.x * v.x, u.y * v.y, u.z * v.z); }       /* :97 */
static inline v3 v_sdiv(v3 v, double t) { return V(v.x / t, v.y / t, v.z / t); }        /* :102 */
static inline double v_len2(v3 v) { return v.x * v.x + v.y * v.y + v.z * v.z; }         /* :112 */
static inline double v_len(v3 v) { return sqrt(v_len2(v)); }                             /* :107 */
static inline v3 v_unit(v3 v) { double l = v_len(v); return V(v.x / l, v.y / l, v.z / l); } /* :117 */
static inline v3 v_neg(v3 v) { return V(-v.x, -v.y, -v.z); }                             /* :123 */
static inline int v_near_zero(v3 v) {                                                    /* :128 */
    const double s = 1e-8;
    return fabs(v.x) < s && fabs(v.y) < s && fabs(v.z) < s;
}
static inline v3 v_reflect(v3 v, v3 n) { return v_sub(v, v_smul(n, 2 * v_dot(v, n))); } /* :134 */
static inline double go_min(double a, double b) { return a < b ? a : b; } /* finite, non-zero-sign-sensitive use only */
static inline v3 v_refract(v3 uv, v3 n, double eta) {                                    /* :140 */
    double cosTheta = go_min(v_dot(v_neg(uv), n), 1.0);
    v3 perp = v_smul(v_add(uv, v_smul(n, cosTheta)), eta);
    v3 par = v_smul(n, -sqrt(fabs(1.0 - v_len2(perp))));
    return v_add(perp, par);
}

/* tcolor.LinearToSrgb (third-party; call site ray/vec3.go:175-177; pinned by
 * ray/vec3_test.go:264-289): clamp, sRGB OETF, x255, round. */
EXPORT uint8_t oracle_linear_to_srgb(double x) {
    if (!(x > 0)) return 0;
    if (x >= 1) return 255;
    double s = x <= 0.0031308 ? 12.92 * x : 1.055 * go_pow_frac(x, 1 / 2.4) - 0.055;
    return (uint8_t)round(255 * s);
}

/* ------------------------------------------------------------------ */
/* Flat scene / camera / params (same flat layout the C-ABI uses)        */
/* ------------------------------------------------------------------ */
enum { MAT_LAMBERTIAN = 0, MAT_METAL = 1, MAT_DIELECTRIC = 2 };

typedef struct {
    int32_t n;
    const double *cx, *cy, *cz, *r;
    const uint8_t *kind;
    const double *params; /* n x 4: albedo rgb + fuzz | refidx,0,0,0 */
    double bg_a[3], bg_b[3];
} oracle_scene;

typedef struct { /* Camera after Initialize (ray/camera.go:9-39) */
    double position[3], pixel00[3], pixel_x[3], pixel_y[3], defocus_u[3], defocus_v[3];
    double aperture, focus_distance, focal_length;
} oracle_camera;

typedef struct { /* user-facing camera fields, before Initialize */
    double position[3], look_at[3], up[3];
    double vfov, focal_length, focus_distance, aperture;
} oracle_camera_in;

typedef struct {
    int32_t width, height, spp, max_depth;
    double ray_radius;
    uint64_t seed;
    int32_t num_workers;  /* reference fan-out (ray/tracer.go:85-116) */
    int32_t stream_mode;  /* 0 = reference chunk streams, 1 = per-sample streams */
    int32_t fma_mode;     /* 0 = strict (Go/amd64), 1 = fused intersection (matches GPU "fma" mode) */
    int32_t threads;      /* OS threads for stream_mode 1 (0 => num_workers) */
} oracle_params;

typedef struct { uint64_t paths, segments, sphere_tests, rng_draws, max_depth_hits; } oracle_stats;

typedef struct { v3 o, d; } ray_t;
typedef struct { v3 p, n; double t; int id; int front; } hit_t;
typedef struct { const oracle_scene *sc; int fma_mode; oracle_stats st; } ctx_t;

static inline v3 A3(const double *p) { return V(p[0], p[1], p[2]); }

/* Sphere.Hit (ray/objects.go:81-104), strict order. */
static inline int sphere_hit(const oracle_scene *sc, int i, const ray_t *r, double tmin, double tmax, hit_t *hr, int fma_mode) {
    v3 c = V(sc->cx[i], sc->cy[i], sc->cz[i]);
    double rad = sc->r[i];
    v3 oc = v_sub(c, r->o);
    double a = v_len2(r->d);
    double h, cc, disc;
    if (!fma_mode) {
        h = v_dot(r->d, oc);
        cc = v_len2(oc) - rad * rad;
        disc = h * h - a * cc;
    } else { /* the fused form the GPU "fma" mode uses: 11 FP64 ops */
        double r2 = rad * rad;
        h = fma(r->d.z, oc.z, fma(r->d.y, oc.y, r->d.x * oc.x));
        cc = fma(oc.z, oc.z, fma(oc.y, oc.y, fma(oc.x, oc.x, -r2)));
        disc = fma(h, h, -(a * cc));
    }
    if (disc < 0) return 0;
    double sq = sqrt(disc);
    double root = (h - sq) / a;
    if (!(root > tmin && root < tmax)) {
        root = (h + sq) / a;
        if (!(root > tmin && root < tmax)) return 0;
    }
    hr->p = v_add(r->o, v_smul(r->d, root)); /* Ray.At, ray/ray.go:23 */
    hr->t = root;
    v3 on = v_sdiv(v_sub(hr->p, c), rad);
    hr->front = v_dot(r->d, on) < 0; /* SetFaceNormal, ray/objects.go:19-26 */
    hr->n = hr->front ? on : v_neg(on);
    hr->id = i;
    return 1;
}

/* Scene.Hit (ray/objects.go:37-46): slice order, strictly-closer wins. */
static int scene_hit(ctx_t *cx, const ray_t *r, double tmin, double tmax, hit_t *hr) {
    int any = 0;
    double closest = tmax;
    const oracle_scene *sc = cx->sc;
    for (int i = 0; i < sc->n; i++) {
        if (sphere_hit(sc, i, r, tmin, closest, hr, cx->fma_mode)) { any = 1; closest = hr->t; }
    }
    cx->st.segments++;
    cx->st.sphere_tests += (uint64_t)sc->n;
    return any;
}

/* AmbientLight.Hit (ray/objects.go:68-73). */
static v3 background(const oracle_scene *sc, const ray_t *r) {
    v3 u = v_unit(r->d);
    double a = 0.5 * (u.y + 1.0);
    return v_add(v_smul(A3(sc->bg_a), 1.0 - a), v_smul(A3(sc->bg_b), a));
}

/* Reflectance (ray/materials.go:66-71). */
static double reflectance(double cosine, double ri) {
    double r0 = (1 - ri) / (1 + ri);
    r0 *= r0;
    return r0 + (1 - r0) * go_pow5(1 - cosine);
}
EXPORT double oracle_reflectance(double cosine, double ri) { return reflectance(cosine, ri); }

/* Material.Scatter (ray/materials.go:13-64). Returns 1 if scattered. */
static int scatter(const oracle_scene *sc, rng_t *rng, const ray_t *rin, const hit_t *hr, v3 *att, ray_t *out) {
    const double *p = sc->params + 4 * hr->id;
    switch (sc->kind[hr->id]) {
    case MAT_LAMBERTIAN: {
        double ux, uy, uz;
        rng_unit_vector(rng, &ux, &uy, &uz);
        v3 dir = v_add(hr->n, V(ux, uy, uz));
        if (v_near_zero(dir)) dir = hr->n;
        out->o = hr->p; out->d = dir;
        *att = V(p[0], p[1], p[2]);
        return 1;
    }
    case MAT_METAL: {
        v3 refl = v_reflect(v_unit(rin->d), hr->n);
        if (p[3] > 0.0) {
            double ux, uy, uz;
            rng_unit_vector(rng, &ux, &uy, &uz);
            refl = v_add(refl, v_smul(V(ux, uy, uz), p[3]));
        }
        out->o = hr->p; out->d = refl;
        if (v_dot(refl, hr->n) > 0) { *att = V(p[0], p[1], p[2]); return 1; }
        return 0;
    }
    default: {
        double ri = p[0];
        double ratio = hr->front ? 1.0 / ri : ri;
        v3 ud = v_unit(rin->d);
        double cosTheta = go_min(v_dot(v_neg(ud), hr->n), 1.0);
        double sinTheta = sqrt(1.0 - cosTheta * cosTheta);
        int cannot = ratio * sinTheta > 1.0;
        v3 dir;
        if (cannot || reflectance(cosTheta, ratio) > rng_f64(rng)) dir = v_reflect(ud, hr->n);
        else dir = v_refract(ud, hr->n, ratio);
        out->o = hr->p; out->d = dir;
        *att = V(1.0, 1.0, 1.0);
        return 1;
    }
    }
}

/* Scene.RayColor (ray/objects.go:49-62): recursive, product applied on unwind. */
static v3 ray_color(ctx_t *cx, rng_t *rng, const ray_t *r, int depth) {
    if (depth <= 0) { cx->st.max_depth_hits++; return V(0, 0, 0); }
    hit_t hr;
    if (scene_hit(cx, r, 1e-6, INFINITY, &hr)) {
        v3 att; ray_t sc;
        if (scatter(cx->sc, rng, r, &hr, &att, &sc)) return v_mul(att, ray_color(cx, rng, &sc, depth - 1));
        return V(0, 0, 0);
    }
    return background(cx->sc, r);
}

/* Camera.GetRay (ray/camera.go:113-142). */
static ray_t get_ray(const oracle_camera *c, rng_t *rng, double px, double py, double ox, double oy) {
    v3 pos = A3(c->position);
    v3 sample = v_add(v_add(A3(c->pixel00), v_smul(A3(c->pixel_x), px + ox)), v_smul(A3(c->pixel_y), py + oy));
    ray_t r;
    r.o = pos;
    r.d = v_sub(sample, pos);
    if (c->aperture > 0) {
        double dx, dy;
        rng_in_disc(rng, 1.0, &dx, &dy);
        v3 offset = v_add(v_smul(A3(c->defocus_u), dx), v_smul(A3(c->defocus_v), dy));
        double focusTime = c->focus_distance / c->focal_length;
        v3 focusPoint = v_add(pos, v_smul(r.d, focusTime));
        r.o = v_add(pos, offset);
        r.d = v_sub(focusPoint, r.o);
    }
    return r;
}

/* One sample of one pixel: ray/tracer.go:133-143 loop body. */
static v3 trace_sample(ctx_t *cx, const oracle_camera *cam, const oracle_params *p, rng_t *rng, int x, int y) {
    double ox = 0.0, oy = 0.0;
    if (p->spp > 1) rng_in_disc(rng, p->ray_radius, &ox, &oy);
    ray_t r = get_ray(cam, rng, (double)x, (double)y, ox, oy);
    cx->st.paths++;
    return ray_color(cx, rng, &r, p->max_depth);
}

typedef struct {
    const oracle_scene *sc; const oracle_camera *cam; const oracle_params *p;
    uint8_t *rgba; size_t stride; double *hdr;
} job_t;

static void store_pixel(const job_t *j, int x, int y, v3 sum) {
    double div = 1.0 / (double)j->p->spp;
    v3 c = v_smul(sum, div);
    if (j->rgba) {
        uint8_t *s = j->rgba + (size_t)y * j->stride + 4 * (size_t)x;
        s[0] = oracle_linear_to_srgb(c.x); s[1] = oracle_linear_to_srgb(c.y); s[2] = oracle_linear_to_srgb(c.z); s[3] = 255;
    }
    if (j->hdr) { double *h = j->hdr + 3 * ((size_t)y * j->p->width + x); h[0] = c.x; h[1] = c.y; h[2] = c.z; }
}

/* Tracer.RenderLines (ray/tracer.go:120-155): ONE stream per call, idx = stream index. */
static void render_lines(const job_t *j, int idx, int y0, int y1, oracle_stats *st) {
    ctx_t cx = {j->sc, j->p->fma_mode, {0, 0, 0, 0, 0}};
    rng_t rng = rng_new_idx((uint64_t)(int64_t)idx, j->p->seed);
    for (int y = y0; y < y1; y++) {
        for (int x = 0; x < (g_xlimit > 0 ? g_xlimit : j->p->width); x++) {
            v3 sum = V(0, 0, 0);
            for (int s = 0; s < j->p->spp; s++) sum = v_add(sum, trace_sample(&cx, j->cam, j->p, &rng, x, y));
            store_pixel(j, x, y, sum);
        }
    }
    cx.st.rng_draws = rng.draws;
    st->paths += cx.st.paths; st->segments += cx.st.segments; st->sphere_tests += cx.st.sphere_tests;
    st->rng_draws += cx.st.rng_draws; st->max_depth_hits += cx.st.max_depth_hits;
}

/* Per-sample stream convention: stream idx = (y*W+x)*spp + s, same seed.
 * A legal use of the reference's own constructor rand.NewIdx; partition-independent. */
static void render_rows_per_sample(const job_t *j, int y0, int y1, oracle_stats *st) {
    ctx_t cx = {j->sc, j->p->fma_mode, {0, 0, 0, 0, 0}};
    uint64_t draws = 0;
    for (int y = y0; y < y1; y++) {
        for (int x = 0; x < j->p->width; x++) {
            v3 sum = V(0, 0, 0);
            for (int s = 0; s < j->p->spp; s++) {
                uint64_t idx = ((uint64_t)y * (uint64_t)j->p->width + (uint64_t)x) * (uint64_t)j->p->spp + (uint64_t)s;
                rng_t rng = rng_new_idx(idx, j->p->seed);
                sum = v_add(sum, trace_sample(&cx, j->cam, j->p, &rng, x, y));
                draws += rng.draws;
            }
            store_pixel(j, x, y, sum);
        }
    }
    st->paths += cx.st.paths; st->segments += cx.st.segments; st->sphere_tests += cx.st.sphere_tests;
    st->rng_draws += draws; st->max_depth_hits += cx.st.max_depth_hits;
}

typedef struct {
    const job_t *job; int chunk, height, mode; int next; pthread_mutex_t mu; oracle_stats st;
} pool_t;

static void *worker(void *arg) {
    pool_t *pl = (pool_t *)arg;
    oracle_stats st = {0, 0, 0, 0, 0};
    for (;;) {
        pthread_mutex_lock(&pl->mu);
        int y = pl->next;
        pl->next += pl->chunk;
        pthread_mutex_unlock(&pl->mu);
        if (y >= pl->height) break;
        int y1 = y + pl->chunk < pl->height ? y + pl->chunk : pl->height;
        if (pl->mode == 0) render_lines(pl->job, y, y, y1, &st);
        else render_rows_per_sample(pl->job, y, y1, &st);
    }
    pthread_mutex_lock(&pl->mu);
    pl->st.paths += st.paths; pl->st.segments += st.segments; pl->st.sphere_tests += st.sphere_tests;
    pl->st.rng_draws += st.rng_draws; pl->st.max_depth_hits += st.max_depth_hits;
    pthread_mutex_unlock(&pl->mu);
    return NULL;
}

/* Tracer.Render fan-out (ray/tracer.go:85-116). Defaults (tracer.go:49-78) are applied by
 * the caller (host mirror); this takes already-defaulted params and an initialised camera. */
EXPORT int oracle_render(const oracle_scene *sc, const oracle_camera *cam, const oracle_params *p,
                         uint8_t *rgba, size_t stride, double *hdr, oracle_stats *stats) {
    if (!sc || !cam || !p || p->width <= 0 || p->height <= 0 || p->spp <= 0) return -1;
    job_t job = {sc, cam, p, rgba, stride, hdr};
    oracle_stats st = {0, 0, 0, 0, 0};
    int W = p->num_workers > 0 ? p->num_workers : 1;
    if (p->stream_mode == 0 && W == 1) {
        render_lines(&job, 0, 0, p->height, &st); /* tracer.go:87-89 */
    } else {
        int chunk = p->height / (W * 4);
        if (chunk < 4) chunk = 4; /* tracer.go:93 */
        int nthreads = p->stream_mode == 0 ? W : (p->threads > 0 ? p->threads : W);
        pool_t pl;
        pl.job = &job; pl.chunk = chunk; pl.height = p->height; pl.mode = p->stream_mode; pl.next = 0;
        memset(&pl.st, 0, sizeof pl.st);
        pthread_mutex_init(&pl.mu, NULL);
        pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
        for (int i = 0; i < nthreads; i++) pthread_create(&th[i], NULL, worker, &pl);
        for (int i = 0; i < nthreads; i++) pthread_join(th[i], NULL);
        free(th);
        pthread_mutex_destroy(&pl.mu);
        st = pl.st;
    }
    if (stats) *stats = st;
    return 0;
}

/* Bounded CPU-baseline sample: rows y == row_offset (mod row_step), per-sample streams, `threads`
 * OS threads pulling rows from a shared counter (same work-queue shape as ray/tracer.go:93-115). */
typedef struct { const job_t *job; int step, next, height; pthread_mutex_t mu; oracle_stats st; } spool_t;
static void *sample_worker(void *arg) {
    spool_t *pl = (spool_t *)arg;
    oracle_stats st = {0, 0, 0, 0, 0};
    for (;;) {
        pthread_mutex_lock(&pl->mu);
        int y = pl->next;
        pl->next += pl->step;
        pthread_mutex_unlock(&pl->mu);
        if (y >= pl->height) break;
        render_rows_per_sample(pl->job, y, y + 1, &st);
    }
    pthread_mutex_lock(&pl->mu);
    pl->st.paths += st.paths; pl->st.segments += st.segments; pl->st.sphere_tests += st.sphere_tests;
    pl->st.rng_draws += st.rng_draws; pl->st.max_depth_hits += st.max_depth_hits;
    pthread_mutex_unlock(&pl->mu);
    return NULL;
}
EXPORT int oracle_render_sampled_rows(const oracle_scene *sc, const oracle_camera *cam, const oracle_params *p,
                                      int row_step, int row_offset, int threads,
                                      uint8_t *rgba, size_t stride, double *hdr, oracle_stats *stats) {
    if (!sc || !cam || !p || row_step <= 0 || threads <= 0) return -1;
    job_t job = {sc, cam, p, rgba, stride, hdr};
    spool_t pl;
    pl.job = &job; pl.step = row_step; pl.next = row_offset; pl.height = p->height;
    memset(&pl.st, 0, sizeof pl.st);
    pthread_mutex_init(&pl.mu, NULL);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    for (int i = 0; i < threads; i++) pthread_create(&th[i], NULL, sample_worker, &pl);
    for (int i = 0; i < threads; i++) pthread_join(th[i], NULL);
    free(th);
    pthread_mutex_destroy(&pl.mu);
    if (stats) *stats = pl.st;
    return 0;
}

/* Sample subsets (per-sample streams): raw colour sums (colorSum of ray/tracer.go:143, before the 1/N) over the samples
 * s = s_off + j*s_stride, j < s_count, of rows [y0,y1), row-major from row y0. accumulate != 0 continues the sums already
 * in `sums` (same running sum as the one-shot loop when the slices are consecutive). Checker for tray_render's sums modes. */
EXPORT int oracle_sample_sums(const oracle_scene *sc, const oracle_camera *cam, const oracle_params *p, int y0, int y1,
                              int s_off, int s_stride, int s_count, int accumulate, double *sums, oracle_stats *stats) {
    if (!sc || !cam || !p || !sums || s_stride < 1 || s_count < 0 || s_off < 0) return -1;
    ctx_t cx = {sc, p->fma_mode, {0, 0, 0, 0, 0}};
    for (int y = y0; y < y1; y++)
        for (int x = 0; x < p->width; x++) {
            double *h = sums + 3 * ((size_t)(y - y0) * p->width + x);
            v3 sum = accumulate ? V(h[0], h[1], h[2]) : V(0, 0, 0);
            for (int j = 0; j < s_count; j++) {
                uint64_t s = (uint64_t)s_off + (uint64_t)j * (uint64_t)s_stride;
                uint64_t idx = ((uint64_t)y * (uint64_t)p->width + (uint64_t)x) * (uint64_t)p->spp + s;
                rng_t rng = rng_new_idx(idx, p->seed);
                sum = v_add(sum, trace_sample(&cx, cam, p, &rng, x, y));
            }
            h[0] = sum.x; h[1] = sum.y; h[2] = sum.z;
        }
    if (stats) *stats = cx.st;
    return 0;
}
/* pixel = ToSRGBA(sum * (1/n_samples)) (ray/tracer.go:145-152) */
EXPORT int oracle_resolve_sums(const double *sums, size_t n_pixels, uint64_t n_samples, uint8_t *rgba) {
    if (!sums || !rgba || !n_samples) return -1;
    double div = 1.0 / (double)n_samples;
    for (size_t k = 0; k < n_pixels; k++) {
        v3 c = v_smul(V(sums[3 * k], sums[3 * k + 1], sums[3 * k + 2]), div);
        rgba[4 * k] = oracle_linear_to_srgb(c.x); rgba[4 * k + 1] = oracle_linear_to_srgb(c.y);
        rgba[4 * k + 2] = oracle_linear_to_srgb(c.z); rgba[4 * k + 3] = 255;
    }
    return 0;
}

/* Camera.GetRay (ray/camera.go:113-142) for n consecutive calls on ONE stream rand.NewIdx(idx, seed), like the reference's
 * camera tests do with RandForTests() = rand.New(42) (camera_test.go:11-13). px/py/ox/oy: n values each; out: n x 6 (origin, dir). */
EXPORT int oracle_get_rays(const oracle_camera *cam, uint64_t idx, uint64_t seed, int n, const double *px, const double *py,
                           const double *ox, const double *oy, double *out) {
    if (!cam || n < 0) return -1;
    rng_t rng = rng_new_idx(idx, seed);
    for (int k = 0; k < n; k++) {
        ray_t r = get_ray(cam, &rng, px[k], py[k], ox[k], oy[k]);
        out[6 * k] = r.o.x; out[6 * k + 1] = r.o.y; out[6 * k + 2] = r.o.z;
        out[6 * k + 3] = r.d.x; out[6 * k + 4] = r.d.y; out[6 * k + 5] = r.d.z;
    }
    return 0;
}

/* Probe of the Vec3 helpers the hot path is built from (ray/vec3.go:25-145), so that the reference's own tables
 * (ray/vec3_test.go) can be replayed against this restatement. op: 0 Add 1 Sub 2 Mul 3 SMul(t) 4 SDiv(t) 5 Cross 6 Unit(u)
 * 7 Neg(u) 8 Reflect(u,n=v) 9 Refract(u,n=v,eta=t) 10 Minus(u; v,w) = u-(v+w) (ray/vec3.go:44-55); scalar results in
 * out[3]: 11 Dot 12 Length(u) 13 LengthSquared(u) 14 NearZero(u) 15 Interval{Start=u.x,End=u.y}.Surrounds(t). */
EXPORT int oracle_vec_op(int op, const double *pu, const double *pv, const double *pw, double t, double *out4) {
    v3 u = A3(pu), v = pv ? A3(pv) : V(0, 0, 0), w = pw ? A3(pw) : V(0, 0, 0), r = V(0, 0, 0);
    double s = 0;
    switch (op) {
        case 0: r = v_add(u, v); break;
        case 1: r = v_sub(u, v); break;
        case 2: r = v_mul(u, v); break;
        case 3: r = v_smul(u, t); break;
        case 4: r = v_sdiv(u, t); break;
        case 5: r = v_cross(u, v); break;
        case 6: r = v_unit(u); break;
        case 7: r = v_neg(u); break;
        case 8: r = v_reflect(u, v); break;
        case 9: r = v_refract(u, v, t); break;
        case 10: r = v_sub(u, v_add(v, w)); break;
        case 11: s = v_dot(u, v); break;
        case 12: s = v_len(u); break;
        case 13: s = v_len2(u); break;
        case 14: s = v_near_zero(u); break;
        case 15: s = (t > u.x && t < u.y) ? 1 : 0; break;  /* Interval.Surrounds, ray/vec3.go:198-200: exclusive */
        default: return -1;
    }
    out4[0] = r.x; out4[1] = r.y; out4[2] = r.z; out4[3] = s;
    return 0;
}

/* Tracer.RenderLines(idx, yStart, yEnd, scene) (ray/tracer.go:120): rows outside stay untouched. */
EXPORT int oracle_render_lines(const oracle_scene *sc, const oracle_camera *cam, const oracle_params *p,
                               int idx, int y0, int y1, uint8_t *rgba, size_t stride, double *hdr, oracle_stats *stats) {
    if (!sc || !cam || !p) return -1;
    job_t job = {sc, cam, p, rgba, stride, hdr};
    oracle_stats st = {0, 0, 0, 0, 0};
    if (p->stream_mode == 0) render_lines(&job, idx, y0, y1, &st);
    else render_rows_per_sample(&job, y0, y1, &st);
    if (stats) *stats = st;
    return 0;
}

/* RNG-independent first hit of the pixel-centre pinhole primary ray (GetRay with offsets 0
 * and the aperture branch skipped) through Scene.Hit with FrontEpsilon. id = -1 on miss. */
EXPORT int oracle_first_hit(const oracle_scene *sc, const oracle_camera *cam, int width, int height, int fma_mode,
                            int32_t *id, double *t, double *normal, uint8_t *front) {
    ctx_t cx = {sc, fma_mode, {0, 0, 0, 0, 0}};
    oracle_camera c = *cam;
    c.aperture = 0;
    for (int y = 0; y < height; y++)
        for (int x = 0; x < width; x++) {
            ray_t r = get_ray(&c, NULL, (double)x, (double)y, 0.0, 0.0);
            hit_t hr;
            size_t k = (size_t)y * width + x;
            if (scene_hit(&cx, &r, 1e-6, INFINITY, &hr)) {
                id[k] = hr.id; t[k] = hr.t; normal[3 * k] = hr.n.x; normal[3 * k + 1] = hr.n.y; normal[3 * k + 2] = hr.n.z;
                front[k] = (uint8_t)hr.front;
            } else {
                id[k] = -1; t[k] = INFINITY; normal[3 * k] = normal[3 * k + 1] = normal[3 * k + 2] = 0; front[k] = 0;
            }
        }
    return 0;
}

/* Camera.Initialize (ray/camera.go:43-105). tan_mode 0 = Go math.Tan restated, 1 = libm tan. */
EXPORT void oracle_camera_init(const oracle_camera_in *in, int width, int height, int tan_mode, oracle_camera *out) {
    v3 zero = V(0, 0, 0);
    v3 pos = A3(in->position), look = A3(in->look_at), up = A3(in->up);
    double fl = in->focal_length, fov = in->vfov, fd = in->focus_distance;
    if (fl == 0) fl = 1.0;
    if (fov == 0) fov = 90.0;
    if (up.x == zero.x && up.y == zero.y && up.z == zero.z) up = V(0, 1, 0);
    if (fd == 0) fd = fl;
    if (pos.x == 0 && pos.y == 0 && pos.z == 0 && look.x == 0 && look.y == 0 && look.z == 0) look = V(0, 0, -1);
    v3 view = v_sub(pos, look);
    if (v_near_zero(view)) view = V(0, 0, 1);
    v3 w = v_unit(view);
    v3 u = v_unit(v_cross(up, w));
    v3 v = v_cross(w, u);
    double defocusRadius = in->aperture / 2;
    v3 dU = v_smul(u, defocusRadius), dV = v_smul(v, defocusRadius);
    double theta = fov * GO_DEG2RAD;
    double vh = 2.0 * fl * (tan_mode ? tan(theta / 2.0) : go_tan(theta / 2.0));
    double aspect = (double)width / (double)height;
    double vw = aspect * vh;
    v3 hor = v_smul(u, vw), ver = v_smul(v, -vh);
    v3 px = v_sdiv(hor, (double)width), py = v_sdiv(ver, (double)height);
    /* Position.Minus(a,b,c) = Position - ((a+b)+c)   (ray/vec3.go:44-55) */
    v3 ul = v_sub(pos, v_add(v_add(v_smul(w, fl), v_smul(hor, 0.5)), v_smul(ver, 0.5)));
    v3 p00 = v_add(ul, v_smul(v_add(px, py), 0.5));
    double *o;
    o = out->position; o[0] = pos.x; o[1] = pos.y; o[2] = pos.z;
    o = out->pixel00; o[0] = p00.x; o[1] = p00.y; o[2] = p00.z;
    o = out->pixel_x; o[0] = px.x; o[1] = px.y; o[2] = px.z;
    o = out->pixel_y; o[0] = py.x; o[1] = py.y; o[2] = py.z;
    o = out->defocus_u; o[0] = dU.x; o[1] = dU.y; o[2] = dU.z;
    o = out->defocus_v; o[0] = dV.x; o[1] = dV.y; o[2] = dV.z;
    out->aperture = in->aperture; out->focus_distance = fd; out->focal_length = fl;
}

/* RichScene (ray/objects.go:132-175), grid half-width generalised (11 in the reference;
 * 50 for BASELINE config 4). Returns the object count; arrays must hold (2*half)^2+4. */
EXPORT int oracle_rich_scene(uint64_t seed, int half, double *cx, double *cy, double *cz, double *r,
                             uint8_t *kind, double *params) {
    rng_t rng = rng_new_idx(0, seed);
    int n = 0;
#define PUSH(X, Y, Z, R, K, P0, P1, P2, P3) do { cx[n] = X; cy[n] = Y; cz[n] = Z; r[n] = R; kind[n] = K; \
        params[4*n] = P0; params[4*n+1] = P1; params[4*n+2] = P2; params[4*n+3] = P3; n++; } while (0)
    PUSH(0, -1000, 0, 1000, MAT_LAMBERTIAN, 0.5, 0.5, 0.5, 0);
    for (int a = -half; a < half; a++)
        for (int b = -half; b < half; b++) {
            double choose = rng_f64(&rng);
            double x = (double)a + 0.9 * rng_f64(&rng);
            double z = (double)b + 0.9 * rng_f64(&rng);
            v3 center = V(x, 0.2, z);
            if (v_len(v_sub(center, V(4, 0.2, 0))) > 0.9) {
                if (choose < 0.8) {
                    double a0 = rng_f64(&rng), a1 = rng_f64(&rng), a2 = rng_f64(&rng);
                    double b0 = rng_f64(&rng), b1 = rng_f64(&rng), b2 = rng_f64(&rng);
                    PUSH(x, 0.2, z, 0.2, MAT_LAMBERTIAN, a0 * b0, a1 * b1, a2 * b2, 0);
                } else if (choose < 0.95) {
                    double a0 = 0.5 + (1.0 - 0.5) * rng_f64(&rng);
                    double a1 = 0.5 + (1.0 - 0.5) * rng_f64(&rng);
                    double a2 = 0.5 + (1.0 - 0.5) * rng_f64(&rng);
                    double fuzz = rng_f64(&rng) * 0.5;
                    PUSH(x, 0.2, z, 0.2, MAT_METAL, a0, a1, a2, fuzz);
                } else {
                    PUSH(x, 0.2, z, 0.2, MAT_DIELECTRIC, 1.5, 0, 0, 0);
                }
            }
        }
    PUSH(0, 1, 0, 1.0, MAT_DIELECTRIC, 1.5, 0, 0, 0);
    PUSH(-4, 1, 0, 1.0, MAT_LAMBERTIAN, 0.4, 0.2, 0.1, 0);
    PUSH(4, 1, 0, 1.0, MAT_METAL, 0.7, 0.6, 0.5, 0.0);
#undef PUSH
    return n;
}

/* ---- stream dumps for known-answer / device-RNG parity tests ---- */
EXPORT void oracle_rng_u64(uint64_t idx, uint64_t seed, int n, uint64_t *out) {
    rng_t r = rng_new_idx(idx, seed);
    for (int i = 0; i < n; i++) out[i] = rng_u64(&r);
}
EXPORT void oracle_rng_f64(uint64_t idx, uint64_t seed, int n, double *out) {
    rng_t r = rng_new_idx(idx, seed);
    for (int i = 0; i < n; i++) out[i] = rng_f64(&r);
}
EXPORT void oracle_rng_norm(uint64_t idx, uint64_t seed, int n, double *out) {
    rng_t r = rng_new_idx(idx, seed);
    for (int i = 0; i < n; i++) out[i] = rng_norm(&r);
}
EXPORT void oracle_rng_unit_vectors(uint64_t idx, uint64_t seed, int n, double *out) {
    rng_t r = rng_new_idx(idx, seed);
    for (int i = 0; i < n; i++) rng_unit_vector(&r, out + 3 * i, out + 3 * i + 1, out + 3 * i + 2);
}
EXPORT void oracle_rng_in_disc(uint64_t idx, uint64_t seed, double radius, int n, double *out) {
    rng_t r = rng_new_idx(idx, seed);
    for (int i = 0; i < n; i++) rng_in_disc(&r, radius, out + 2 * i, out + 2 * i + 1);
}
EXPORT double oracle_go_log(double x) { return go_log(x); }
EXPORT double oracle_go_exp(double x) { return go_exp(x); }
EXPORT double oracle_go_sin(double x) { return go_sin(x); }
EXPORT double oracle_go_cos(double x) { return go_cos(x); }
EXPORT double oracle_go_tan(double x) { return go_tan(x); }

/* Single-object / single-op probes so the reference's unit-test tables
 * can be re-expressed against the oracle. */
EXPORT int oracle_sphere_hit(const double c[3], double radius, const double o[3], const double d[3],
                             double tmin, double tmax, int fma_mode, double *t, double *point, double *normal, int *front) {
    double cx = c[0], cy = c[1], cz = c[2], r = radius; uint8_t k = 0; double prm[4] = {0, 0, 0, 0};
    oracle_scene sc; memset(&sc, 0, sizeof sc);
    sc.n = 1; sc.cx = &cx; sc.cy = &cy; sc.cz = &cz; sc.r = &r; sc.kind = &k; sc.params = prm;
    ray_t ry = {A3(o), A3(d)};
    hit_t hr;
    if (!sphere_hit(&sc, 0, &ry, tmin, tmax, &hr, fma_mode)) return 0;
    *t = hr.t; point[0] = hr.p.x; point[1] = hr.p.y; point[2] = hr.p.z;
    normal[0] = hr.n.x; normal[1] = hr.n.y; normal[2] = hr.n.z; *front = hr.front;
    return 1;
}

/* Material.Scatter probe: returns did_scatter; writes attenuation and scattered ray. */
EXPORT int oracle_scatter(int kind, const double prm[4], uint64_t idx, uint64_t seed,
                          const double rin_o[3], const double rin_d[3], const double point[3], const double normal[3],
                          int front, double *att, double *out_o, double *out_d, uint64_t *draws) {
    double cx = 0, cy = 0, cz = 0, r = 1; uint8_t k = (uint8_t)kind;
    oracle_scene sc; memset(&sc, 0, sizeof sc);
    sc.n = 1; sc.cx = &cx; sc.cy = &cy; sc.cz = &cz; sc.r = &r; sc.kind = &k; sc.params = prm;
    rng_t rng = rng_new_idx(idx, seed);
    ray_t rin = {A3(rin_o), A3(rin_d)}, out = {V(0, 0, 0), V(0, 0, 0)};
    hit_t hr; hr.p = A3(point); hr.n = A3(normal); hr.t = 0; hr.id = 0; hr.front = front;
    v3 a = V(0, 0, 0);
    int did = scatter(&sc, &rng, &rin, &hr, &a, &out);
    att[0] = a.x; att[1] = a.y; att[2] = a.z;
    out_o[0] = out.o.x; out_o[1] = out.o.y; out_o[2] = out.o.z;
    out_d[0] = out.d.x; out_d[1] = out.d.y; out_d[2] = out.d.z;
    if (draws) *draws = rng.draws;
    return did;
}

/* RayColor probe for a single ray with its own stream (objects_test.go:241-320 style). */
EXPORT void oracle_ray_color(const oracle_scene *sc, const double o[3], const double d[3], int depth,
                             uint64_t idx, uint64_t seed, int fma_mode, double *rgb) {
    ctx_t cx = {sc, fma_mode, {0, 0, 0, 0, 0}};
    rng_t rng = rng_new_idx(idx, seed);
    ray_t r = {A3(o), A3(d)};
    v3 c = ray_color(&cx, &rng, &r, depth);
    rgb[0] = c.x; rgb[1] = c.y; rgb[2] = c.z;
}

/* ------------------------------------------------------------------ */
/* "Next" row 8(f)-1: the interactive path after Render (main.go:119-130) */
/* ------------------------------------------------------------------ */
/* golang.org/x/image/draw v0.35.0 (go.mod:11, NOT in the reference tree): BiLinear.Scale(dst, dst.Bounds(), src,
 * src.Bounds(), draw.Over, nil) for *image.RGBA -> freshly allocated *image.RGBA (call site main.go:127).
 * Restated from the published algorithm (Kernel scaler: separable triangle filter whose support is widened by the
 * scale factor when shrinking; float64 intermediates in [0,1]; 16-bit premultiplied composition). PARITY UNPINNED:
 * no golden vector for it exists in the reference (both mains are excluded from its tests). */
typedef struct { int i, j; double inv, inv_ffff; double center, arg_scale; } bl_src;

static bl_src bl_source(int x, int dw, int sw) {
    double scale = (double)sw / (double)dw;
    double half = 1.0, arg = 1.0;
    if (scale > 1) { half *= scale; arg = 1 / scale; }
    double center = ((double)x + 0.5) * scale - 0.5;
    int i = (int)floor(center - half);
    if (i < 0) i = 0;
    int j = (int)ceil(center + half);
    if (j > sw) { j = sw; if (j < i) j = i; }
    double total = 0.0;
    for (int c = i; c < j; c++) {
        double t = fabs((center - (double)c) * arg);
        if (t >= 1.0) continue;
        double w = 1.0 - t;
        if (w == 0) continue;
        total += w;
    }
    bl_src s;
    s.i = i; s.j = j; s.center = center; s.arg_scale = arg;
    s.inv = 1 / total;
    s.inv_ffff = s.inv / 0xffff;
    return s;
}
static inline double bl_weight(const bl_src *s, int c) {
    double t = fabs((s->center - (double)c) * s->arg_scale);
    if (t >= 1.0) return 0.0;
    return 1.0 - t;
}
static uint16_t bl_ftou(double f) {
    int32_t i = (int32_t)(0xffff * f + 0.5);
    if (i > 0xffff) return 0xffff;
    if (i > 0) return (uint16_t)i;
    return 0;
}

/* draw.NearestNeighbor.Scale(dst, dst.Bounds(), src, src.Bounds(), draw.Over, nil) onto a fresh RGBA (main.go:124-125;
 * golang.org/x/image/draw nnInterpolator, restated: third-party, parity unpinned): integer source coordinates
 * sx = (2*dx+1)*sw/(2*dw), sy likewise; 16-bit premultiplied Over onto a zeroed destination. */
EXPORT int oracle_nn_scale(const uint8_t *src, int sw, int sh, size_t sstride, int dw, int dh, uint8_t *dst /* dw*dh*4, zeroed */) {
    if (sw <= 0 || sh <= 0 || dw <= 0 || dh <= 0) return -1;
    uint64_t dw2 = 2 * (uint64_t)dw, dh2 = 2 * (uint64_t)dh;
    for (int dy = 0; dy < dh; dy++) {
        uint64_t sy = (2 * (uint64_t)dy + 1) * (uint64_t)sh / dh2;
        for (int dx = 0; dx < dw; dx++) {
            uint64_t sx = (2 * (uint64_t)dx + 1) * (uint64_t)sw / dw2;
            const uint8_t *px = src + (size_t)sy * sstride + 4 * (size_t)sx;
            uint32_t pa = (uint32_t)px[3] * 0x101, a1 = (0xffff - pa) * 0x101;
            uint8_t *d = dst + 4 * ((size_t)dy * dw + dx);
            for (int k = 0; k < 4; k++) d[k] = (uint8_t)(((uint32_t)d[k] * a1 / 0xffff + (uint32_t)px[k] * 0x101) >> 8);
        }
    }
    return 0;
}

EXPORT int oracle_bilinear_scale(const uint8_t *src, int sw, int sh, size_t sstride, int dw, int dh, uint8_t *dst /* dw*dh*4, zeroed */) {
    if (sw <= 0 || sh <= 0 || dw <= 0 || dh <= 0) return -1;
    double *tmp = (double *)malloc(sizeof(double) * 4 * (size_t)dw * (size_t)sh);
    for (int y = 0; y < sh; y++)
        for (int x = 0; x < dw; x++) {
            bl_src s = bl_source(x, dw, sw);
            double p[4] = {0, 0, 0, 0};
            for (int c = s.i; c < s.j; c++) {
                double w = bl_weight(&s, c);
                if (w == 0) continue;
                const uint8_t *px = src + (size_t)y * sstride + 4 * (size_t)c;
                for (int k = 0; k < 4; k++) p[k] += (double)((uint32_t)px[k] * 0x101) * w;
            }
            double *t = tmp + 4 * ((size_t)y * dw + x);
            for (int k = 0; k < 4; k++) t[k] = p[k] * s.inv_ffff;
        }
    for (int x = 0; x < dw; x++)
        for (int y = 0; y < dh; y++) {
            bl_src s = bl_source(y, dh, sh);
            double p[4] = {0, 0, 0, 0};
            for (int c = s.i; c < s.j; c++) {
                double w = bl_weight(&s, c);
                if (w == 0) continue;
                const double *t = tmp + 4 * ((size_t)c * dw + x);
                for (int k = 0; k < 4; k++) p[k] += t[k] * w;
            }
            for (int k = 0; k < 3; k++) if (p[k] > p[3]) p[k] = p[3];
            uint32_t q[4];
            for (int k = 0; k < 4; k++) q[k] = bl_ftou(p[k] * s.inv);
            uint32_t a1 = (0xffff - q[3]) * 0x101;
            uint8_t *d = dst + 4 * ((size_t)y * dw + x);
            for (int k = 0; k < 4; k++) d[k] = (uint8_t)(((uint32_t)d[k] * a1 / 0xffff + q[k]) >> 8);
        }
    free(tmp);
    return 0;
}

/* Half-block truecolor frame (the data format ansipixels.ShowScaledImage emits, main.go:130: two image rows per
 * terminal row, U+2584 with the upper pixel as background and the lower pixel as foreground). The exact byte
 * stream of ansipixels is not in the reference tree; this library defines a fixed-width record so that cells can be
 * written independently: ESC[48;2;RRR;GGG;BBBm ESC[38;2;RRR;GGG;BBBm E2 96 84 (41 bytes), rows end ESC[0m LF (5). */
#define ANSI_CELL 41
#define ANSI_EOL 5
static void ansi_cell(uint8_t *o, const uint8_t *top, const uint8_t *bot) {
    static const char hdr[2][8] = {"\x1b[48;2;", "\x1b[38;2;"};
    const uint8_t *px[2] = {top, bot};
    for (int h = 0; h < 2; h++) {
        memcpy(o, hdr[h], 7); o += 7;
        for (int k = 0; k < 3; k++) {
            unsigned v = px[h][k];
            *o++ = (uint8_t)('0' + v / 100); *o++ = (uint8_t)('0' + (v / 10) % 10); *o++ = (uint8_t)('0' + v % 10);
            *o++ = k < 2 ? ';' : 'm';
        }
    }
    *o++ = 0xE2; *o++ = 0x96; *o++ = 0x84;
}
EXPORT size_t oracle_ansi_halfblocks(const uint8_t *img, int w, int h2 /* even */, uint8_t *out) {
    size_t n = 0;
    for (int r = 0; r < h2 / 2; r++) {
        for (int x = 0; x < w; x++, n += ANSI_CELL)
            ansi_cell(out + n, img + 4 * ((size_t)(2 * r) * w + x), img + 4 * ((size_t)(2 * r + 1) * w + x));
        memcpy(out + n, "\x1b[0m\n", ANSI_EOL);
        n += ANSI_EOL;
    }
    return n;
}
