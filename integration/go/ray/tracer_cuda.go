//go:build cuda

// tracer_cuda.go -- cgo shim that routes (*Tracer).Render / RenderLines to libtraycuda.so (B200, sm_100a).
//
// Drop this file into the reference's ray/ package and build with `-tags cuda` (CGO_ENABLED=1). The pure-Go
// bodies of Render/RenderLines in ray/tracer.go move behind `//go:build !cuda` (see INTEGRATION.md for the
// 12-line diff). Everything else in the package -- New, the Tracer fields, Camera.Initialize, Scene, Sphere,
// the materials, RichScene -- is unchanged, so tray and benchmark keep their flags and behaviour.
//
// NOT COMPILED IN THE BUILD CONTAINER (no Go toolchain there): it is kept deliberately thin -- marshalling
// only -- and the identical C ABI is exercised from Python (ctypes) and C++ (tray_b200/host) by the test-suite.
package ray

/*
#cgo CFLAGS: -I${SRCDIR}/../../../include
#cgo LDFLAGS: -L${SRCDIR}/../../../tray_b200 -ltraycuda -Wl,-rpath,${SRCDIR}/../../../tray_b200
#include <stdlib.h>
#include "tray_cuda.h"
*/
import "C"

import (
	"crypto/rand"
	"encoding/binary"
	"fmt"
	"image"
	"os"
	"runtime"
	"strconv"
	"sync"
	"time"
	"unsafe"
)

// Backend knobs (additive; zero values keep the drop-in defaults).
var (
	// CUDADevices lists the CUDA ordinals one Tracer context uses (nil = device 0). TRAY_GPUS=N sets 0..N-1.
	CUDADevices []int
	// CUDAReferenceStreams reproduces the reference's sequential per-chunk RNG streams (one warp per chunk,
	// honours NumWorkers; slow by construction). Default: per-sample streams, independent of NumWorkers.
	CUDAReferenceStreams = os.Getenv("TRAY_STREAMS") == "reference"
	// CUDAFusedFP64 selects the fused-multiply-add discriminant (faster, not bit-identical to Go/amd64).
	CUDAFusedFP64 = os.Getenv("TRAY_FP64") == "fma"
	// CUDAFloat32 selects the float32 fast path (1.35x the default on a B200, ~61 dB PSNR against it; no parity claim).
	CUDAFloat32 = os.Getenv("TRAY_FP32") == "1"
	// CUDASampleSplit splits samples instead of tiles across devices (partial sums reduced over NVLink).
	CUDASampleSplit = os.Getenv("TRAY_SPLIT") == "samples"
)

type cudaState struct {
	mu  sync.Mutex
	ctx *C.tray_ctx
}

var cuda cudaState

func cudaContext() *C.tray_ctx {
	if cuda.ctx != nil {
		return cuda.ctx
	}
	devs := CUDADevices
	if n, err := strconv.Atoi(os.Getenv("TRAY_GPUS")); err == nil && n > 0 && devs == nil {
		for i := 0; i < n; i++ {
			devs = append(devs, i)
		}
	}
	var ctx *C.tray_ctx
	var rc C.int
	if len(devs) == 0 {
		rc = C.tray_init(nil, 0, &ctx)
	} else {
		cd := make([]C.int, len(devs))
		for i, d := range devs {
			cd[i] = C.int(d)
		}
		rc = C.tray_init(&cd[0], C.int(len(cd)), &ctx)
	}
	if rc != 0 {
		panic(fmt.Sprintf("tray: CUDA backend unavailable: %s", C.GoString(C.tray_last_error(nil))))
	}
	cuda.ctx = ctx
	return ctx
}

// flatScene is the SoA marshalling of Scene.Objects: flat numeric slices only (cgo pointer rules).
type flatScene struct {
	cx, cy, cz, r []float64
	kind          []uint8
	params        []float64
}

func (f *flatScene) walk(objs []Hittable) error {
	for _, o := range objs {
		switch s := o.(type) {
		case *Scene: // Scene satisfies Hittable (objects.go:28-46): flattened in place, order preserved
			if err := f.walk(s.Objects); err != nil {
				return err
			}
		case *Sphere:
			var k uint8
			var p [4]float64
			switch m := s.Mat.(type) {
			case Lambertian:
				k, p = C.TRAY_MAT_LAMBERTIAN, [4]float64{m.Albedo.x, m.Albedo.y, m.Albedo.z, 0}
			case Metal:
				k, p = C.TRAY_MAT_METAL, [4]float64{m.Albedo.x, m.Albedo.y, m.Albedo.z, m.Fuzz}
			case Dielectric:
				k, p = C.TRAY_MAT_DIELECTRIC, [4]float64{m.RefIdx, 0, 0, 0}
			case *Lambertian: // value receivers: pointers satisfy Material as well
				k, p = C.TRAY_MAT_LAMBERTIAN, [4]float64{m.Albedo.x, m.Albedo.y, m.Albedo.z, 0}
			case *Metal:
				k, p = C.TRAY_MAT_METAL, [4]float64{m.Albedo.x, m.Albedo.y, m.Albedo.z, m.Fuzz}
			case *Dielectric:
				k, p = C.TRAY_MAT_DIELECTRIC, [4]float64{m.RefIdx, 0, 0, 0}
			default:
				return fmt.Errorf("tray cuda: unsupported Material %T (no CPU fallback)", s.Mat)
			}
			f.cx = append(f.cx, s.Center.x)
			f.cy = append(f.cy, s.Center.y)
			f.cz = append(f.cz, s.Center.z)
			f.r = append(f.r, s.Radius)
			f.kind = append(f.kind, k)
			f.params = append(f.params, p[:]...)
		default:
			return fmt.Errorf("tray cuda: unsupported Hittable %T (no CPU fallback)", o)
		}
	}
	return nil
}

func v3(dst *[3]C.double, v Vec3) { dst[0], dst[1], dst[2] = C.double(v.x), C.double(v.y), C.double(v.z) }

func ptrF(s []float64) *C.double {
	if len(s) == 0 {
		return nil
	}
	return (*C.double)(unsafe.Pointer(&s[0]))
}

func randomSeed() uint64 {
	var b [8]byte
	_, _ = rand.Read(b[:])
	return binary.LittleEndian.Uint64(b[:]) | 1
}

// renderCUDA renders rows [y0,y1) into t.imageData. idx < 0: Render (chunking from NumWorkers); else RenderLines(idx).
func (t *Tracer) renderCUDA(scene *Scene, y0, y1, idx int) error {
	cuda.mu.Lock()
	defer cuda.mu.Unlock()
	ctx := cudaContext()
	var f flatScene
	if err := f.walk(scene.Objects); err != nil {
		return err
	}
	// tray_scene_desc holds POINTERS to the SoA arrays. cgo only lets C see a Go pointer to memory that itself contains Go
	// pointers when those inner pointers are pinned (cgocheck: "Go pointer to unpinned Go pointer" panics otherwise), so the
	// backing arrays are pinned for the duration of the upload (runtime.Pinner, Go >= 1.21; the module requires go 1.24).
	// The library copies everything before tray_scene_upload returns and keeps no caller pointer.
	var pin runtime.Pinner
	defer pin.Unpin()
	var sd C.tray_scene_desc
	sd.n = C.int32_t(len(f.cx))
	if len(f.cx) > 0 {
		pin.Pin(&f.cx[0])
		pin.Pin(&f.cy[0])
		pin.Pin(&f.cz[0])
		pin.Pin(&f.r[0])
		pin.Pin(&f.params[0])
		pin.Pin(&f.kind[0])
		sd.cx, sd.cy, sd.cz, sd.radius, sd.mat_params = ptrF(f.cx), ptrF(f.cy), ptrF(f.cz), ptrF(f.r), ptrF(f.params)
		sd.mat_kind = (*C.uint8_t)(unsafe.Pointer(&f.kind[0]))
	}
	v3(&sd.bg_a, scene.Background.ColorA)
	v3(&sd.bg_b, scene.Background.ColorB)
	if rc := C.tray_scene_upload(ctx, &sd); rc != 0 {
		return fmt.Errorf("tray cuda: %s", C.GoString(C.tray_last_error(ctx)))
	}
	var cam C.tray_camera
	v3(&cam.position, t.Position)
	v3(&cam.pixel00, t.pixel00)
	v3(&cam.pixel_x, t.pixelXVector)
	v3(&cam.pixel_y, t.pixelYVector)
	v3(&cam.defocus_u, t.defocusDiskU)
	v3(&cam.defocus_v, t.defocusDiskV)
	cam.aperture, cam.focus_distance, cam.focal_length = C.double(t.Aperture), C.double(t.FocusDistance), C.double(t.FocalLength)
	var p C.tray_params
	p.width, p.height = C.int32_t(t.width), C.int32_t(t.height)
	p.spp, p.max_depth, p.ray_radius = C.int32_t(t.NumRaysPerPixel), C.int32_t(t.MaxDepth), C.double(t.RayRadius)
	seed := t.Seed
	if seed == 0 { // "0 means randomized each time" (tracer.go:32)
		seed = randomSeed()
	}
	p.seed = C.uint64_t(seed)
	p.y0, p.y1 = C.int32_t(y0), C.int32_t(y1)
	p.stream_mode = C.TRAY_STREAM_PER_SAMPLE
	if CUDAReferenceStreams {
		p.stream_mode = C.TRAY_STREAM_REFERENCE
	}
	p.num_workers, p.stream_idx = C.int32_t(t.NumWorkers), -1
	if idx >= 0 {
		p.num_workers, p.stream_idx = 0, C.int64_t(idx)
	}
	p.precision = C.TRAY_FP64_STRICT
	if CUDAFusedFP64 {
		p.precision = C.TRAY_FP64_FMA
	}
	if CUDAFloat32 {
		p.precision = C.TRAY_FP32
	}
	if CUDASampleSplit {
		p.split_mode = C.TRAY_SPLIT_SAMPLES
	}
	if cudaSubset.mode != 0 { // RenderProgressive / multi-process sample split: raw colour sums stay on the device
		p.sample_offset, p.sample_stride, p.sample_count = C.int32_t(cudaSubset.offset), C.int32_t(cudaSubset.stride), C.int32_t(cudaSubset.count)
		p.sums_mode = cudaSubset.mode
	}
	switch os.Getenv("TRAY_LAYOUT") { // default: plain megakernel; the other layouts give the same bits
	case "regroup":
		p.layout = C.TRAY_LAYOUT_REGROUP
	case "wavefront":
		p.layout = C.TRAY_LAYOUT_WAVEFRONT
	}
	// ProgressFunc: poll the library from a goroutine; deltas sum to the pixel count (tracer_test.go:172-186).
	done := make(chan struct{})
	var wg sync.WaitGroup
	total, sent := (y1-y0)*t.width, 0
	if t.ProgressFunc != nil {
		wg.Add(1)
		go func() {
			defer wg.Done()
			tick := time.NewTicker(20 * time.Millisecond)
			defer tick.Stop()
			for {
				select {
				case <-done:
					return
				case <-tick.C:
					if cur := min(int(C.tray_progress(ctx)), total); cur > sent {
						t.ProgressFunc(cur - sent)
						sent = cur
					}
				}
			}
		}()
	}
	pix := t.imageData.Pix
	var st C.tray_stats
	rc := C.tray_render(ctx, &cam, &p, (*C.uint8_t)(unsafe.Pointer(&pix[0])), C.size_t(t.imageData.Stride), &st)
	close(done)
	wg.Wait()
	runtime.KeepAlive(f)
	if rc != 0 {
		return fmt.Errorf("tray cuda: %s", C.GoString(C.tray_last_error(ctx)))
	}
	if t.ProgressFunc != nil && total > sent {
		t.ProgressFunc(total - sent)
	}
	return nil
}

// Render performs the ray tracing and returns the resulting image data (same contract as the pure-Go Render).
func (t *Tracer) Render(scene *Scene) *image.RGBA {
	scene = t.prepare(scene) // the defaulting block of tracer.go:49-83, unchanged, factored into a helper
	if err := t.renderCUDA(scene, 0, t.height, -1); err != nil {
		panic(err) // Render has no error return; RenderErr is the additive variant
	}
	return t.imageData
}

// RenderErr is Render with an error return (additive API).
func (t *Tracer) RenderErr(scene *Scene) (*image.RGBA, error) {
	scene = t.prepare(scene)
	err := t.renderCUDA(scene, 0, t.height, -1)
	return t.imageData, err
}

// RenderLines renders rows [yStart,yEnd) only; idx is the stream index in reference-stream mode.
func (t *Tracer) RenderLines(idx, yStart, yEnd int, scene *Scene) {
	if err := t.renderCUDA(scene, yStart, yEnd, idx); err != nil {
		panic(err)
	}
}

// ---- additive API over the rest of the C ABI (the callers either side of Render in main.go / benchmark.go) ----

// EncodePNG returns the PNG file of the last rendered frame, encoded on the GPU from the frame still resident in
// HBM (tray_encode_png). It replaces png.Encode inside SaveImage (main.go:26-36, benchmark/benchmark.go:23-33):
//
//	func SaveImage(rt *ray.Tracer, fname string) error { b, err := rt.EncodePNG(); if err != nil { return err }; return os.WriteFile(fname, b, 0o644) }
//
// Decoding the file gives back exactly img.Pix's R,G,B (8-bit truecolour, what image/png writes for an opaque RGBA).
func (t *Tracer) EncodePNG() ([]byte, error) {
	cuda.mu.Lock()
	defer cuda.mu.Unlock()
	ctx := cudaContext()
	buf := make([]byte, int(C.tray_png_bound(C.int32_t(t.width), C.int32_t(t.height))))
	var n C.size_t
	if rc := C.tray_encode_png(ctx, (*C.uint8_t)(unsafe.Pointer(&buf[0])), C.size_t(len(buf)), &n, nil); rc != 0 {
		return nil, fmt.Errorf("tray cuda: %s", C.GoString(C.tray_last_error(ctx)))
	}
	return buf[:int(n)], nil
}

// Present is the tail of tray's OnResize (main.go:119-130) on the GPU: the last frame scaled to cols x 2*rows pixels
// (draw.BiLinear, or draw.NearestNeighbor when the frame is smaller than the terminal: -s < 1) and emitted as the
// half-block truecolor frame for a cols x rows terminal. Only the ANSI bytes cross PCIe; write them to the terminal
// in place of ap.ShowScaledImage(resized).
func (t *Tracer) Present(cols, rows int) ([]byte, error) {
	cuda.mu.Lock()
	defer cuda.mu.Unlock()
	ctx := cudaContext()
	out := make([]byte, rows*(cols*41+5))
	var n C.size_t
	rc := C.tray_present(ctx, C.int32_t(cols), C.int32_t(2*rows), nil, (*C.uint8_t)(unsafe.Pointer(&out[0])), C.size_t(len(out)), &n, nil)
	if rc != 0 {
		return nil, fmt.Errorf("tray cuda: %s", C.GoString(C.tray_last_error(ctx)))
	}
	return out[:int(n)], nil
}

// RenderProgressive delivers the same frame as Render as a sequence of refinements (README "WIP: navigation in the
// world"): after every `slice` rays per pixel, show(raysDone, img) is called with the mean over the rays so far. The
// scene and the per-pixel colour sums stay on the device and each slice continues the sum in sample order
// (tray_params.sums_mode), so the last image is bit-identical to Render's. show may return false to stop early
// (a key was pressed: re-render from the new camera).
func (t *Tracer) RenderProgressive(scene *Scene, slice int, show func(raysDone int, img *image.RGBA) bool) error {
	scene = t.prepare(scene)
	seed := t.Seed
	if seed == 0 {
		seed = randomSeed()
	}
	saved := t.Seed
	t.Seed = seed
	defer func() { t.Seed = saved }()
	for done := 0; done < t.NumRaysPerPixel; {
		k := min(slice, t.NumRaysPerPixel-done)
		mode := C.int32_t(C.TRAY_SUMS_ACCUMULATE)
		if done == 0 {
			mode = C.TRAY_SUMS_OVERWRITE
		}
		if err := t.renderSubset(scene, done, 1, k, mode); err != nil { // renderCUDA with p.sample_* / p.sums_mode set
			return err
		}
		done += k
		cuda.mu.Lock()
		ctx := cudaContext()
		pix := t.imageData.Pix
		rc := C.tray_resolve_sums(ctx, C.uint64_t(done), (*C.uint8_t)(unsafe.Pointer(&pix[0])), C.size_t(t.imageData.Stride))
		var msg string
		if rc != 0 {
			msg = C.GoString(C.tray_last_error(ctx))
		}
		cuda.mu.Unlock()
		if rc != 0 {
			return fmt.Errorf("tray cuda: %s", msg)
		}
		if show != nil && !show(done, t.imageData) {
			return nil
		}
	}
	return nil
}

// subset carries the sample-subset fields of tray_params for one renderCUDA call (zero = all samples, classic render).
type subset struct {
	offset, stride, count int
	mode                  C.int32_t
}

var cudaSubset subset // read by renderCUDA when filling tray_params: p.sample_offset/stride/count, p.sums_mode

func (t *Tracer) renderSubset(scene *Scene, offset, stride, count int, mode C.int32_t) error {
	cudaSubset = subset{offset, stride, count, mode}
	defer func() { cudaSubset = subset{} }()
	return t.renderCUDA(scene, 0, t.height, -1)
}
