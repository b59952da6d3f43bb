// tray.hpp -- C++ host layer above the C ABI (include/tray_cuda.h), mirroring the reference's Go package `ray`
// for the hot path: same type and method names, argument meaning, defaults and side effects
// (fortio/tray ray/tracer.go, ray/camera.go, ray/objects.go, ray/materials.go, ray/vec3.go).
// Scene construction, Camera::Initialize and Tracer defaulting run here on the host exactly like the Go code;
// everything per pixel runs in libtraycuda.so. There is no CPU rendering path.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <memory>
#include <random>
#include <stdexcept>
#include <string>
#include <variant>
#include <vector>

#include "../../include/tray_cuda.h"

namespace fortio_rand {  // host side of fortio.org/rand (go.mod:9): what scene construction needs
class Rand {
   public:
    // rand.NewIdx(idx, seed) == Go math/rand/v2 NewPCG(uint64(idx), seed); seed 0 randomizes (main.go:47)
    Rand(uint64_t seed, uint64_t idx = 0) {
        if (seed == 0) { std::random_device rd; seed = ((uint64_t)rd() << 32 | rd()) | 1; idx = rd(); }
        hi_ = idx; lo_ = seed;
    }
    uint64_t Uint64() {
        const unsigned __int128 MUL = ((unsigned __int128)2549297995355413924ULL << 64) | 4865540595714422341ULL;
        const unsigned __int128 INC = ((unsigned __int128)6364136223846793005ULL << 64) | 1442695040888963407ULL;
        unsigned __int128 s = (((unsigned __int128)hi_ << 64) | lo_) * MUL + INC;
        hi_ = (uint64_t)(s >> 64); lo_ = (uint64_t)s;
        uint64_t hi = hi_;
        hi ^= hi >> 32; hi *= 0xda942042e4dd58b5ULL; hi ^= hi >> 48; hi *= (lo_ | 1);
        return hi;
    }
    double Float64() { return (double)(Uint64() << 11 >> 11) / 9007199254740992.0; }
    double Float64Range(double a, double b) { return a + (b - a) * Float64(); }
   private:
    uint64_t hi_, lo_;
};
inline Rand New(uint64_t seed) { return Rand(seed, 0); }
inline Rand NewIdx(int idx, uint64_t seed) { return Rand(seed, (uint64_t)(int64_t)idx); }
}  // namespace fortio_rand

namespace ray {

struct Vec3 { double x = 0, y = 0, z = 0; };
using ColorF = Vec3;
inline Vec3 Add(Vec3 u, Vec3 v) { return {v.x + u.x, v.y + u.y, v.z + u.z}; }
inline Vec3 Sub(Vec3 u, Vec3 v) { return {u.x - v.x, u.y - v.y, u.z - v.z}; }
inline Vec3 SMul(Vec3 v, double t) { return {v.x * t, v.y * t, v.z * t}; }
inline Vec3 SDiv(Vec3 v, double t) { return {v.x / t, v.y / t, v.z / t}; }
inline Vec3 Mul(Vec3 u, Vec3 v) { return {u.x * v.x, u.y * v.y, u.z * v.z}; }
inline Vec3 Cross(Vec3 u, Vec3 v) { return {u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x}; }
inline double LengthSquared(Vec3 v) { return v.x * v.x + v.y * v.y + v.z * v.z; }
inline double Length(Vec3 v) { return std::sqrt(LengthSquared(v)); }
inline Vec3 Unit(Vec3 v) { double l = Length(v); return {v.x / l, v.y / l, v.z / l}; }
inline bool NearZero(Vec3 v) { const double s = 1e-8; return std::fabs(v.x) < s && std::fabs(v.y) < s && std::fabs(v.z) < s; }
inline bool IsZero(Vec3 v) { return v.x == 0 && v.y == 0 && v.z == 0; }

struct Lambertian { ColorF Albedo; };
struct Metal { ColorF Albedo; double Fuzz = 0; };
struct Dielectric { double RefIdx = 1; };
using Material = std::variant<Lambertian, Metal, Dielectric>;
struct Sphere { Vec3 Center; double Radius = 0; Material Mat; };
struct AmbientLight { ColorF ColorA, ColorB; };
inline AmbientLight DefaultBackground() { return {{1.0, 1.0, 1.0}, {0.4, 0.65, 1.0}}; }  // ray/objects.go:106-110

struct Scene {  // Scene{Objects, Background}; only spheres exist on the GPU path
    std::vector<Sphere> Objects;
    AmbientLight Background;
};

inline Scene DefaultScene() {  // ray/objects.go:112-130
    Scene s;
    s.Objects = {{{0, 0, -1.2}, 0.5, Lambertian{{0.1, 0.2, 0.5}}}, {{0, -100.5, -1}, 100, Lambertian{{0.7, 0.8, 0.1}}},
                 {{-1.0, 0, -1}, 0.5, Dielectric{1.5}},            {{-1.0, 0, -1}, 0.4, Dielectric{1.0 / 1.5}},
                 {{1.0, 0, -1}, 0.5, Metal{{1, .8, .8}, 0.05}}};
    s.Background = DefaultBackground();
    return s;
}

inline Scene RichScene(fortio_rand::Rand& rng, int half = 11) {  // ray/objects.go:132-175 (half=50: BASELINE config 4)
    Scene w;
    w.Objects.push_back({{0, -1000, 0}, 1000, Lambertian{{0.5, 0.5, 0.5}}});
    for (int a = -half; a < half; a++)
        for (int b = -half; b < half; b++) {
            double chooseMat = rng.Float64();
            double cx = (double)a + 0.9 * rng.Float64();
            double cz = (double)b + 0.9 * rng.Float64();
            Vec3 center{cx, 0.2, cz};
            if (Length(Sub(center, Vec3{4, 0.2, 0})) > 0.9) {
                if (chooseMat < 0.8) {
                    Vec3 r1{rng.Float64(), rng.Float64(), rng.Float64()};
                    Vec3 r2{rng.Float64(), rng.Float64(), rng.Float64()};
                    w.Objects.push_back({center, 0.2, Lambertian{Mul(r1, r2)}});
                } else if (chooseMat < 0.95) {
                    double r = rng.Float64Range(0.5, 1.0), g = rng.Float64Range(0.5, 1.0), bl = rng.Float64Range(0.5, 1.0);
                    double fuzz = rng.Float64() * 0.5;
                    w.Objects.push_back({center, 0.2, Metal{{r, g, bl}, fuzz}});
                } else {
                    w.Objects.push_back({center, 0.2, Dielectric{1.5}});
                }
            }
        }
    w.Objects.push_back({{0, 1, 0}, 1.0, Dielectric{1.5}});
    w.Objects.push_back({{-4, 1, 0}, 1.0, Lambertian{{0.4, 0.2, 0.1}}});
    w.Objects.push_back({{4, 1, 0}, 1.0, Metal{{0.7, 0.6, 0.5}, 0.0}});
    return w;
}

// Go math.Tan (src/math/tan.go: pure-Go Cephes form, no assembly on amd64 / arm64) for |x| < 2^29 -- what Camera.Initialize
// calls (ray/camera.go:85). Only + - * / on doubles (build with -ffp-contract=off); libm's tan may differ in the last ulp.
inline double GoTan(double x) {
    const double PI4A = 7.85398125648498535156e-1, PI4B = 3.77489470793079817668e-8, PI4C = 2.69515142907905952645e-15;
    const double P0 = -1.30936939181383777646e4, P1 = 1.15351664838587416140e6, P2 = -1.79565251976484877988e7;
    const double Q1 = 1.36812963470692954678e4, Q2 = -1.32089234440210967447e6, Q3 = 2.50083801823357915839e7, Q4 = -5.38695755929454629881e7;
    if (x == 0 || std::isnan(x)) return x;
    if (std::isinf(x)) return std::nan("");
    const bool sign = x < 0;
    if (sign) x = -x;
    uint64_t j = (uint64_t)(x * 0x1.45f306dc9c883p+0);  // 4/Pi, folded like Go folds the constant
    double y = (double)j;
    if (j & 1) { j++; y++; }
    const double z = ((x - y * PI4A) - y * PI4B) - y * PI4C, zz = z * z;
    if (zz > 1e-14) y = z + z * (zz * (((P0 * zz) + P1) * zz + P2) / ((((zz + Q1) * zz + Q2) * zz + Q3) * zz + Q4));
    else y = z;
    if (j & 2) y = -1 / y;
    return sign ? -y : y;
}

struct Camera {  // ray/camera.go:9-39
    Vec3 Position, LookAt, Up;
    double VerticalFoV = 0, FocalLength = 0, FocusDistance = 0, Aperture = 0;
    Vec3 pixel00, pixelXVector, pixelYVector, defocusDiskU, defocusDiskV;

    void Initialize(int width, int height) {  // ray/camera.go:43-105
        if (FocalLength == 0) FocalLength = 1.0;
        if (VerticalFoV == 0) VerticalFoV = 90.0;
        if (IsZero(Up)) Up = {0, 1, 0};
        if (FocusDistance == 0) FocusDistance = FocalLength;
        if (IsZero(Position) && IsZero(LookAt)) LookAt = {0, 0, -1};
        Vec3 view = Sub(Position, LookAt);
        if (NearZero(view)) view = {0, 0, 1};
        Vec3 w = Unit(view), u = Unit(Cross(Up, w)), v = Cross(w, u);
        double defocusRadius = Aperture / 2;
        defocusDiskU = SMul(u, defocusRadius);
        defocusDiskV = SMul(v, defocusRadius);
        double theta = VerticalFoV * 0x1.1df46a2529d39p-6;  // Go constant math.Pi/180.0, exactly rounded
        double viewportHeight = 2.0 * FocalLength * GoTan(theta / 2.0);  // math.Tan as Go computes it (Cephes form), not libm's
        double aspectRatio = (double)width / (double)height;
        double viewportWidth = aspectRatio * viewportHeight;
        Vec3 horizontal = SMul(u, viewportWidth), vertical = SMul(v, -viewportHeight);
        pixelXVector = SDiv(horizontal, (double)width);
        pixelYVector = SDiv(vertical, (double)height);
        Vec3 upperLeft = Sub(Position, Add(Add(SMul(w, FocalLength), SMul(horizontal, 0.5)), SMul(vertical, 0.5)));
        pixel00 = Add(upperLeft, SMul(Add(pixelXVector, pixelYVector), 0.5));
    }
};
inline Camera RichSceneCamera() {  // ray/camera.go:144-154
    Camera c;
    c.Position = {13, 2, 3}; c.LookAt = {0, 0, 0}; c.Up = {0, 1, 0};
    c.VerticalFoV = 20.0; c.Aperture = 0.1; c.FocalLength = 10.0; c.FocusDistance = 10.0;
    return c;
}

struct RGBA {  // image.RGBA: Pix, Stride, bounds
    int W = 0, H = 0, Stride = 0;
    std::vector<uint8_t> Pix;
};

class Tracer : public Camera {  // ray/tracer.go:25-35
   public:
    int MaxDepth = 0, NumRaysPerPixel = 0, NumWorkers = 0;
    double RayRadius = 0;
    uint64_t Seed = 0;
    // additive backend knobs
    int StreamMode = TRAY_STREAM_PER_SAMPLE, Precision = TRAY_FP64_STRICT, SplitMode = TRAY_SPLIT_TILES;
    std::vector<int> Devices;  // CUDA ordinals; empty = device 0
    tray_stats Stats{};

    Tracer(int width, int height) : width_(width), height_(height) {  // ray.New
        imageData_.W = width; imageData_.H = height; imageData_.Stride = 4 * width;
        imageData_.Pix.assign((size_t)4 * width * height, 0);
    }
    ~Tracer() { if (ctx_) tray_destroy(ctx_); }
    Tracer(const Tracer&) = delete;
    Tracer& operator=(const Tracer&) = delete;
    void SetCamera(const Camera& c) { static_cast<Camera&>(*this) = c; }

    // (*Tracer).Render (ray/tracer.go:48-118): returns the tracer's own image. scene may be null.
    RGBA& Render(Scene* scene) {
        Scene local;
        if (!scene) {  // tracer.go:49-61
            local = DefaultScene(); scene = &local;
            Position = {-2, 2, 1}; LookAt = {0, 0, -1}; VerticalFoV = 20.0; Aperture = .1;
            FocusDistance = Length(Sub(Position, LookAt));
        }
        if (IsZero(scene->Background.ColorA) && IsZero(scene->Background.ColorB)) scene->Background = DefaultBackground();
        if (MaxDepth <= 0) MaxDepth = 10;
        if (NumRaysPerPixel <= 0) NumRaysPerPixel = 1;
        if (RayRadius <= 0) RayRadius = 0.5;
        if (NumWorkers <= 0) NumWorkers = 1;
        Initialize(width_, height_);
        run(*scene, 0, height_, -1, NumWorkers);
        return imageData_;
    }
    // (*Tracer).RenderLines (ray/tracer.go:120-155)
    void RenderLines(int idx, int yStart, int yEnd, Scene& scene) { run(scene, yStart, yEnd, idx, 0); }
    RGBA& Image() { return imageData_; }
    uint64_t Progress() const { return ctx_ ? tray_progress(ctx_) : 0; }
    // The tail of tray's OnResize (main.go:119-130) on the device: the last frame scaled to cols x 2*rows pixels and
    // emitted as the half-block truecolor frame of a cols x rows terminal (write it to the terminal as is).
    std::string Present(int cols, int rows, double* device_ms = nullptr) {
        if (!ctx_) throw std::runtime_error("Present: nothing rendered");
        std::string out((size_t)rows * ((size_t)cols * 41 + 5), '\0');
        size_t n = 0;
        check(tray_present(ctx_, cols, 2 * rows, nullptr, reinterpret_cast<uint8_t*>(&out[0]), out.size(), &n, device_ms));
        out.resize(n);
        return out;
    }
    int Layout = TRAY_LAYOUT_AUTO;
    // SaveImage(img, fname) (main.go:26-36, benchmark/benchmark.go:23-33): the PNG file of the last frame, encoded on the
    // device from the frame still resident in HBM. Returns the device time of the encode in ms.
    double SaveImage(const std::string& fname) {
        if (!ctx_) throw std::runtime_error("SaveImage: nothing rendered");
        std::vector<uint8_t> png(tray_png_bound(width_, height_));
        size_t n = 0;
        double ms = 0;
        check(tray_encode_png(ctx_, png.data(), png.size(), &n, &ms));
        FILE* f = fopen(fname.c_str(), "wb");
        if (!f) throw std::runtime_error("could not create " + fname);
        bool ok = fwrite(png.data(), 1, n, f) == n;
        ok = (fclose(f) == 0) && ok;
        if (!ok) throw std::runtime_error("could not write " + fname);
        return ms;
    }

   private:
    void check(int rc) {
        if (rc != TRAY_OK) throw std::runtime_error(std::string("libtraycuda: ") + tray_last_error(ctx_));
    }
    void run(Scene& scene, int y0, int y1, int64_t stream_idx, int workers) {
        if (!ctx_) check(tray_init(Devices.empty() ? nullptr : Devices.data(), (int)Devices.size(), &ctx_));
        size_t n = scene.Objects.size();
        std::vector<double> cx(n), cy(n), cz(n), r(n), prm(4 * n);
        std::vector<uint8_t> kind(n);
        for (size_t i = 0; i < n; i++) {
            const Sphere& s = scene.Objects[i];
            cx[i] = s.Center.x; cy[i] = s.Center.y; cz[i] = s.Center.z; r[i] = s.Radius;
            if (auto* l = std::get_if<Lambertian>(&s.Mat)) { kind[i] = TRAY_MAT_LAMBERTIAN; prm[4*i] = l->Albedo.x; prm[4*i+1] = l->Albedo.y; prm[4*i+2] = l->Albedo.z; }
            else if (auto* m = std::get_if<Metal>(&s.Mat)) { kind[i] = TRAY_MAT_METAL; prm[4*i] = m->Albedo.x; prm[4*i+1] = m->Albedo.y; prm[4*i+2] = m->Albedo.z; prm[4*i+3] = m->Fuzz; }
            else { kind[i] = TRAY_MAT_DIELECTRIC; prm[4*i] = std::get<Dielectric>(s.Mat).RefIdx; }
        }
        tray_scene_desc d{};
        d.n = (int32_t)n; d.cx = cx.data(); d.cy = cy.data(); d.cz = cz.data(); d.radius = r.data();
        d.mat_kind = kind.data(); d.mat_params = prm.data();
        const AmbientLight& bg = scene.Background;
        d.bg_a[0] = bg.ColorA.x; d.bg_a[1] = bg.ColorA.y; d.bg_a[2] = bg.ColorA.z;
        d.bg_b[0] = bg.ColorB.x; d.bg_b[1] = bg.ColorB.y; d.bg_b[2] = bg.ColorB.z;
        check(tray_scene_upload(ctx_, &d));
        tray_camera c{};
        auto put = [](double* o, Vec3 v) { o[0] = v.x; o[1] = v.y; o[2] = v.z; };
        put(c.position, Position); put(c.pixel00, pixel00); put(c.pixel_x, pixelXVector); put(c.pixel_y, pixelYVector);
        put(c.defocus_u, defocusDiskU); put(c.defocus_v, defocusDiskV);
        c.aperture = Aperture; c.focus_distance = FocusDistance; c.focal_length = FocalLength;
        tray_params p{};
        p.width = width_; p.height = height_; p.spp = NumRaysPerPixel; p.max_depth = MaxDepth; p.ray_radius = RayRadius;
        uint64_t seed = Seed;
        if (seed == 0) { std::random_device rd; seed = (((uint64_t)rd() << 32) | rd()) | 1; }  // tracer.go:32
        p.seed = seed; p.y0 = y0; p.y1 = y1; p.stream_mode = StreamMode; p.num_workers = workers; p.stream_idx = stream_idx;
        p.precision = Precision; p.split_mode = SplitMode; p.layout = Layout;
        check(tray_render(ctx_, &c, &p, imageData_.Pix.data(), (size_t)imageData_.Stride, &Stats));
    }
    int width_, height_;
    RGBA imageData_;
    tray_ctx* ctx_ = nullptr;
};

inline std::unique_ptr<Tracer> New(int width, int height) { return std::make_unique<Tracer>(width, height); }

}  // namespace ray
