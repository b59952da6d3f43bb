/*
 * tray_cuda.h -- C ABI of libtraycuda.so, the B200 (sm_100a) path-tracing backend for fortio/tray.
 *
 * This is the drop-in boundary: the Go package `ray` keeps its API
 * (ray.New, Tracer fields, (*Tracer).Render / RenderLines, Scene/Sphere/materials, Camera) and
 * its cgo shim binds exactly these entry points (see INTEGRATION.md for the stub). Plain
 * pointers and sizes only; the library copies every input before returning and keeps no
 * caller pointer (cgo pointer rules). All functions return 0 on success or a negative
 * TRAY_E_* code; the message is available from tray_last_error().
 *
 * Reference interfaces replaced (paths relative to the reference repo):
 *   tray_render        <- (*Tracer).Render        ray/tracer.go:48-118   (fan-out + accumulate + sRGB store)
 *                      <- (*Tracer).RenderLines   ray/tracer.go:120-155  (rows [y0,y1), stream index idx)
 *   tray_scene_upload  <- Scene{Objects,Background} walk: Scene.Hit ray/objects.go:37-46,
 *                         Sphere ray/objects.go:75-79, Lambertian/Metal/Dielectric ray/materials.go:9-44
 *   tray_camera        <- Camera after Initialize  ray/camera.go:9-39,43-105 (Initialize stays on the host)
 *   tray_first_hit     <- Scene.Hit + Sphere.Hit   ray/objects.go:37-46,81-104 (parity probe, RNG-free)
 *   tray_resolve_sums  <- the tail of RenderLines   ray/tracer.go:145-152 (colorSum * 1/N, ToSRGBA, Pix store)
 *   tray_present       <- draw.BiLinear/NearestNeighbor.Scale + ap.ShowScaledImage   main.go:119-130 (tray's OnResize tail)
 *   tray_configure     <- (no reference counterpart: where the closest-hit BVH of a large scene is built)
 *   tray_cluster_tables <- Scene.Hit           ray/objects.go:37-46 (the structure that replaces the linear scan; host-only probe)
 *   tray_encode_png    <- SaveImage / png.Encode    main.go:26-36, benchmark/benchmark.go:23-33
 *   tray_upload_frame  <- (no reference counterpart: an image the context did not render, for present / PNG / tests)
 *   tray_progress      <- Tracer.ProgressFunc      ray/tracer.go:30,126-128 (poll; deltas sum to w*h)
 *   tray_rng_dump      <- fortio.org/rand streams  ray/tracer.go:121, ray/rand.go:10-32 (parity probe)
 *   tray_arith_probe   <- float64 `/` and math.Sqrt as Go compiles them (DIVSD / SQRTSD): Unit, SDiv ray/vec3.go:60-75,
 *                         Sphere.Hit roots and normal ray/objects.go:90-100 (parity probe of the device's IEEE routines)
 */
#ifndef TRAY_CUDA_H
#define TRAY_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TRAY_ABI_VERSION 3

#if defined(__GNUC__)
#define TRAY_API __attribute__((visibility("default")))
#else
#define TRAY_API
#endif

/* error codes */
#define TRAY_OK 0
#define TRAY_E_INVALID (-1)   /* bad argument */
#define TRAY_E_CUDA (-2)      /* CUDA runtime error (message has the cudaError string) */
#define TRAY_E_NO_SCENE (-3)  /* tray_render before tray_scene_upload */
#define TRAY_E_UNSUPPORTED (-4)
#define TRAY_E_NO_DEVICE (-5) /* no usable sm_100 device: there is NO CPU fallback */

/* material kinds (Material implementations, ray/materials.go) */
#define TRAY_MAT_LAMBERTIAN 0 /* params: albedo r,g,b, unused        */
#define TRAY_MAT_METAL 1      /* params: albedo r,g,b, fuzz          */
#define TRAY_MAT_DIELECTRIC 2 /* params: ref_idx, unused x3          */

/* RNG stream conventions */
#define TRAY_STREAM_REFERENCE 0  /* one sequential stream per reference chunk, exactly ray/tracer.go:87-121
                                    (conformance mode: one warp per chunk, slow by construction) */
#define TRAY_STREAM_PER_SAMPLE 1 /* stream = rand.NewIdx((y*W+x)*spp+s, seed): throughput mode, result
                                    independent of partitioning and GPU count */

/* arithmetic modes */
#define TRAY_FP64_FMA 0    /* float64; Sphere.Hit discriminant uses fused multiply-add (11 FP64 ops/test) */
#define TRAY_FP64_STRICT 1 /* float64; every op rounded separately = Go/amd64 semantics. DEFAULT. The sphere loop runs an
                              fp32 conservative pre-filter (packed FFMA2) that skips a test only when it PROVES the
                              strict fp64 test returns false; survivors get the exact 17-op fp64 test: same bits. */
#define TRAY_FP32 2        /* float32 fast path (reported separately, PSNR vs fp64) */
#define TRAY_FP64_STRICT_BRUTE 3 /* as STRICT but every test in fp64 (no pre-filter): the pure FP64-pipe kernel */

/* closest-hit structure. Results are identical for all of them (ties resolve to the lowest index). */
#define TRAY_ACCEL_AUTO 0  /* cluster boxes up to 32768 spheres (two levels in shared memory up to 2048, three levels in global memory above), BVH beyond */
#define TRAY_ACCEL_BRUTE 1 /* linear scan of the sphere table (the reference's Scene.Hit order) */
#define TRAY_ACCEL_BVH 2   /* small BVH (<= 4 spheres per leaf), built at upload on the host (median split) or on the device (LBVH) */
#define TRAY_ACCEL_CLUSTER 3 /* two-level boxes over chunks of 8 spheres, tested warp-wide with a conservative fp32 slab test; only the
                                chunks some lane of the warp may hit run the pair pre-filter (strict fp64 / fp32 modes; <= 32768 spheres, with a third
                                level of boxes -- one per 512 slots -- and the tables in global memory above 2048) */

/* divergence layout of the trace kernel. Results are identical for all of them. */
#define TRAY_LAYOUT_AUTO 0
#define TRAY_LAYOUT_PLAIN 1   /* megakernel, a path stays in its lane from generation to termination */
#define TRAY_LAYOUT_REGROUP 2 /* megakernel with per-material regrouping: after Scene.Hit the paths of a CTA are sorted through
                                 shared memory by what happens next (miss | Lambertian | Metal | Dielectric), so scatter,
                                 generators, unwind and regeneration run on mostly homogeneous warps */

#define TRAY_LAYOUT_WAVEFRONT 3 /* path state in HBM, one bounce = intersect | shade (per-material queues) | regenerate kernels
                                  (strict fp64 linear scan only; other modes fall back to the regroup megakernel) */

/* multi-GPU partitioning inside one context */
#define TRAY_SPLIT_TILES 0   /* interleaved row bands; device-to-host gather only */
#define TRAY_SPLIT_SAMPLES 1 /* each GPU renders samples s == g (mod G); partial sums reduced over NVLink */

/* sample-subset modes (tray_params.sums_mode) */
#define TRAY_SUMS_OFF 0        /* classic: all samples, mean, sRGB image */
#define TRAY_SUMS_OVERWRITE 1  /* sums := sum over this call's samples */
#define TRAY_SUMS_ACCUMULATE 2 /* sums := previous sums continued with this call's samples (same geometry required) */

typedef struct tray_ctx tray_ctx;

/* Flattened scene: SoA over spheres, in Scene.Objects order (ties resolve to the lowest index). */
typedef struct {
    int32_t n;                 /* number of spheres */
    const double *cx, *cy, *cz; /* centres */
    const double *radius;
    const uint8_t *mat_kind;   /* TRAY_MAT_* */
    const double *mat_params;  /* n x 4 */
    double bg_a[3], bg_b[3];   /* AmbientLight ColorA / ColorB (ray/objects.go:64-73) */
} tray_scene_desc;

/* Camera AFTER Camera.Initialize (ray/camera.go:43-105): derived vectors, computed on the host. */
typedef struct {
    double position[3];
    double pixel00[3], pixel_x[3], pixel_y[3];
    double defocus_u[3], defocus_v[3];
    double aperture, focus_distance, focal_length;
} tray_camera;

typedef struct {
    int32_t width, height;
    int32_t spp;        /* NumRaysPerPixel (>0, already defaulted) */
    int32_t max_depth;  /* MaxDepth (>0, already defaulted) */
    double ray_radius;  /* RayRadius */
    uint64_t seed;      /* non-zero (the host shim draws one when the user passes 0) */
    int32_t y0, y1;     /* rows to render, [y0,y1); Render uses 0,height */
    int32_t stream_mode;  /* TRAY_STREAM_* */
    int32_t num_workers;  /* REFERENCE mode: NumWorkers of the fan-out (ray/tracer.go:85-116);
                             <=0 with stream_idx>=0 means a single RenderLines(stream_idx,y0,y1) stream */
    int64_t stream_idx;   /* REFERENCE mode RenderLines idx; -1 = derive from chunking */
    int32_t precision;    /* TRAY_FP64_FMA | TRAY_FP64_STRICT | TRAY_FP32 */
    int32_t split_mode;   /* TRAY_SPLIT_* (only matters when the context has >1 device) */
    int32_t shard_index, shard_count; /* multi-process tile sharding: this context renders only the
                             row bands b with b % shard_count == shard_index (0,1 or 0,0 = everything);
                             rows of other shards in the output are left untouched */
    int32_t accel;        /* TRAY_ACCEL_*: closest-hit structure (0 = automatic) */
    /* Sample subsets (TRAY_SUMS_* != OFF, per-sample streams only): this call traces the samples
     * s = sample_offset + j*sample_stride, j in [0,sample_count), of every pixel and leaves the RAW colour sums
     * (colorSum of ray/tracer.go:143, before the 1/N) in the context instead of an image; tray_resolve_sums turns
     * them into pixels. Used for (a) multi-process sample split: rank g of G passes offset g, stride G, reduces the
     * sums with NCCL (tray_device_sums exposes the device buffer) and the root resolves; (b) progressive refinement
     * (camera navigation, main.go:143-163): consecutive slices offset k0, stride 1 with TRAY_SUMS_ACCUMULATE continue
     * the sum in sample order, so the image after the last slice is bit-identical to the one-shot render. */
    int32_t sample_offset, sample_stride, sample_count;
    int32_t sums_mode;    /* TRAY_SUMS_* */
    int32_t layout;       /* TRAY_LAYOUT_*: how the megakernel deals with divergence (0 = automatic) */
    int32_t reserved[3];
} tray_params;

typedef struct {
    uint64_t paths;        /* samples traced (w*rows*spp) */
    uint64_t segments;     /* Scene.Hit calls, primary included */
    uint64_t sphere_tests; /* Sphere.Hit evaluations: segments*n for the linear scan; spheres of the scanned chunks (clusters); exact tests (BVH) */
    uint64_t depth_exhausted; /* paths that ran into MaxDepth */
    double kernel_ms;      /* device time of the trace+resolve kernels (CUDA events, max over devices) */
    double total_ms;       /* wall time inside tray_render, copies included */
    int32_t launches;      /* kernels launched by this call */
    int32_t n_devices;
    double trace_kernel_ms; /* device time of the trace kernels alone */
    double trace_launches;  /* how many trace-kernel launches that was (one per pass and device; wavefront: one per bounce stage) */
    double box_tests;       /* TRAY_ACCEL_CLUSTER: conservative box tests evaluated (groups + chunks), summed over rays */
    double bounds_violations; /* bounds-check builds (-DTRAY_BOUNDS_CHECK) only: out-of-range table / list / stack / scratch accesses the
                                 trace kernels refused so far in this process; always 0 in the release build */
} tray_stats;

/* Creates a context on the given CUDA devices (devices==NULL or n_devices<=0: device 0). */
TRAY_API int tray_init(const int *devices, int n_devices, tray_ctx **out);
TRAY_API void tray_destroy(tray_ctx *ctx);
TRAY_API const char *tray_last_error(tray_ctx *ctx); /* ctx may be NULL: last init error */
TRAY_API int tray_abi_version(void);

/* Context options. TRAY_CFG_BVH_BUILD: where tray_scene_upload builds the closest-hit BVH (results are identical for any
 * tree): AUTO = LBVH on the device from 1024 spheres up, host median split below; the device build falls back to the
 * host when its radix tree would be deeper than the traversal stack. tray_query(TRAY_CFG_BVH_BUILD) tells what the last
 * upload did (1 = built on the device). */
#define TRAY_CFG_BVH_BUILD 1
#define TRAY_BVH_BUILD_AUTO 0
#define TRAY_BVH_BUILD_HOST 1
#define TRAY_BVH_BUILD_DEVICE 2
/* TRAY_CFG_INDISC / TRAY_CFG_UNITVEC: which body of fortio.org/rand's Rand.InDisc / Rand.UnitVector the device generators use.
 * Those bodies are not in the reference tree and no reference test pins their values (call sites ray/tracer.go:138,
 * ray/camera.go:128, ray/rand.go:31); 0 = the default restatement, the others are the candidate bodies the oracle also knows
 * (oracle_set_variants): InDisc 1 = polar (angle from the 1st uniform, radius sqrt of the 2nd), 2 = polar the other way round;
 * UnitVector 1 = rejection in the cube (ray/rand.go:50-58), 2 = spherical angles (ray/rand.go:62-69). Every variant is bit-exact
 * against the oracle's; once real Go vectors are available (tools/go_vectors) the matching body is a flag. */
#define TRAY_CFG_INDISC 2
#define TRAY_CFG_UNITVEC 3
TRAY_API int tray_configure(tray_ctx *ctx, int32_t key, int64_t value);
TRAY_API int64_t tray_query(tray_ctx *ctx, int32_t key);

/* Device-free (host code only, no CUDA call): the two-level cluster tables tray_scene_upload stages for TRAY_ACCEL_CLUSTER,
 * for tests and inspection. Layout of the blob (float4 units): pair pre-filter table in slot order (8 float4 per chunk of 8
 * slots) | chunk boxes from meta[1] (three float4 per pair of chunks: centre, half extent) | group boxes from meta[2] | word boxes
 * (one per 8 groups, padded to a multiple of 8 words) right after them | uint16 sphere id per slot from meta[3]. meta = {blob float4s, off_box2, off_box1, off_ids, real groups (multiple of 8), always-groups,
 * chunk mask of the last always-group, spheres the filter cannot bound}; r_out = max |coordinate| the error bounds assume.
 * Copies at most cap_floats floats; blob_out may be NULL (size query). Scene.Hit reference: ray/objects.go:37-46. */
TRAY_API int tray_cluster_tables(const tray_scene_desc *scene, float *blob_out, size_t cap_floats, int32_t *meta, float *r_out);

/* Copies the scene to every device of the context. May be called again to replace the scene. */
TRAY_API int tray_scene_upload(tray_ctx *ctx, const tray_scene_desc *scene);

/* Renders rows [y0,y1) into rgba_out (R,G,B,255 per pixel; row y at rgba_out + y*stride, i.e.
 * rgba_out is &img.Pix[0]). rgba_out may be NULL: the image then stays on the device
 * (fetch with tray_read_image). stats may be NULL. */
TRAY_API int tray_render(tray_ctx *ctx, const tray_camera *cam, const tray_params *params,
                uint8_t *rgba_out, size_t stride, tray_stats *stats);

/* Copies the last rendered image / linear-HDR means (w*h*3 doubles, colorSum*(1/N) before sRGB). */
TRAY_API int tray_read_image(tray_ctx *ctx, uint8_t *rgba_out, size_t stride);
TRAY_API int tray_read_hdr(tray_ctx *ctx, double *hdr_out);

/* Sample-subset renders (tray_params.sums_mode != 0). tray_device_sums returns the DEVICE pointer of the raw colour sums of
 * the last such render (rows_rendered*width*3 doubles, row-major over the rows this context rendered; single-device
 * contexts only; valid until the next render) so that a caller can reduce it across processes in place (NCCL).
 * tray_resolve_sums computes pixel = ToSRGBA(sum * (1/n_samples)) (ray/tracer.go:145-152) for the rows this context
 * rendered; n_samples 0 = the number of samples accumulated by this context so far. The sums stay untouched.
 * The caller must have synchronised any stream of its own that wrote the sums. rgba_out may be NULL. */
TRAY_API int tray_device_sums(tray_ctx *ctx, double **device_ptr, uint64_t *n_doubles);
TRAY_API int tray_resolve_sums(tray_ctx *ctx, uint64_t n_samples, uint8_t *rgba_out, size_t stride);

/* Places an arbitrary RGBA8 frame (row y at rgba + y*stride) on the context's first device as "the last rendered frame", so
 * that tray_present / tray_encode_png / tray_read_image also serve images this context did not render (and so that the
 * tests can feed the encoder adversarial content). */
TRAY_API int tray_upload_frame(tray_ctx *ctx, const uint8_t *rgba, size_t stride, int32_t width, int32_t height);

/* The interactive path after Render (main.go:119-130), on the device: scale the last rendered frame (which must be
 * complete and resident on one device) to cols x rows2 pixels with draw.BiLinear semantics (x/image/draw, draw.Over onto
 * a fresh image) -- or draw.NearestNeighbor when the frame is narrower than cols (tray -s < 1) --, then emit the half-block truecolor frame for a cols x rows2/2 terminal: per cell the fixed-width
 * record ESC[48;2;RRR;GGG;BBBm ESC[38;2;RRR;GGG;BBBm U+2584 (41 bytes), per row ESC[0m LF (5 bytes).
 * rgba_small_out (cols*rows2*4) and ansi_out may each be NULL. ansi_len receives (rows2/2)*(cols*41+5). */
TRAY_API int tray_present(tray_ctx *ctx, int32_t cols, int32_t rows2, uint8_t *rgba_small_out, uint8_t *ansi_out, size_t ansi_cap,
                          size_t *ansi_len, double *device_ms);

/* The -save path (SaveImage -> png.Encode, main.go:26-36, benchmark/benchmark.go:23-33) on the device: encodes the last
 * rendered frame (complete, resident on one device) as an 8-bit truecolour PNG -- what Go writes for an opaque
 * *image.RGBA -- and copies the FILE bytes to png_out. Per-row adaptive filter, dynamic-Huffman deflate blocks, Adler-32
 * and CRC-32 all run on the GPU; decoding the file returns exactly the rendered pixels (format parity = decoded pixels,
 * not file bytes). png_len always receives the file size; png_out may be NULL (size query after encoding) and must hold
 * png_len bytes otherwise (tray_png_bound gives a safe capacity up front). */
TRAY_API size_t tray_png_bound(int32_t width, int32_t height);
TRAY_API int tray_encode_png(tray_ctx *ctx, uint8_t *png_out, size_t cap, size_t *png_len, double *device_ms);

/* RNG-independent parity probe: closest hit of every pixel-centre pinhole primary ray.
 * id -1 / t +Inf on a miss. Arrays are width*height (normal: x3). */
TRAY_API int tray_first_hit(tray_ctx *ctx, const tray_camera *cam, int32_t width, int32_t height, int32_t precision,
                   int32_t *id, double *t, double *normal, uint8_t *front);

/* Parity probe for the device generators: kind 0 Uint64 (as double bit patterns in out64),
 * 1 Float64, 2 NormFloat64, 3 UnitVector (3 per draw), 4 InDisc(radius) (2 per draw). */
TRAY_API int tray_rng_dump(tray_ctx *ctx, int32_t kind, uint64_t idx, uint64_t seed, double radius, int32_t n, double *out);

/* Parity probe for the device's IEEE division / square root routines (one shared reciprocal per divisor):
 * kind 0: (a[i], a[i+1], a[i+2]) / b[i] (3 results per element, indices mod n), 1: a[i] / b[i] through the shared-reciprocal
 * form, 2: sqrt(a[i]), 3: a[i] / b[i] as the compiler emits it. The Go reference gets these from the CPU's IEEE instructions
 * (ray/vec3.go:60-75 Unit / SDiv, ray/objects.go:90-100). */
TRAY_API int tray_arith_probe(tray_ctx *ctx, int32_t kind, const double *a, const double *b, int32_t n, double *out);

/* Parity probe for the on-device sRGB store (ColorF.ToSRGBA, ray/vec3.go:173-180). */
TRAY_API int tray_linear_to_srgb(tray_ctx *ctx, const double *x, int32_t n, uint8_t *out);

/* Pixels completed by the render in flight (or the last one); safe from another thread. */
TRAY_API uint64_t tray_progress(tray_ctx *ctx);

/* Measured FP64 (kind 0: DFMA, 1: DADD+DMUL) / FP32 (kind 2: FFMA) issue-bound peak in TFLOP/s on device 0
 * of the context and the SM clock seen (roofline denominator; bench.py reports it). */
TRAY_API int tray_measure_peak(tray_ctx *ctx, int32_t kind, double *tflops, double *ms);

#ifdef __cplusplus
}
#endif
#endif /* TRAY_CUDA_H */
