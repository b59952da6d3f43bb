// go_vectors -- dumps what pins the B200 backend's oracle to the real Go reference (fortio/tray + fortio.org/rand v1.1.0 +
// fortio.org/terminal's tcolor), for a maintainer with a Go toolchain. The build container of tray-b200 has none, so two
// wrapper bodies (Rand.InDisc, Rand.UnitVector) and the amd64 assembly forms of math.Exp / math.Pow are restated from their
// published algorithms and marked "parity unpinned" in DESIGN.md section 2. This program closes that gap:
//
//	cd <fortio/tray checkout> && cp -r <tray-b200>/tools/go_vectors ./cmd_go_vectors && go run ./cmd_go_vectors > go_vectors.json
//	cp go_vectors.json <tray-b200>/tests/golden/go_vectors.json && python -m pytest tests/test_go_vectors.py
//
// tests/test_go_vectors.py (skipped while the file is absent) compares the C oracle and, on a GPU box, the device
// generators / conformance kernel with every vector, trying each InDisc / UnitVector variant the oracle and the device
// carry (tray_configure(TRAY_CFG_INDISC / TRAY_CFG_UNITVEC)) and reporting which one matches.
package main

import (
	"encoding/json"
	"math"
	mrand "math/rand/v2"
	"os"
	"runtime"

	"fortio.org/rand"
	"fortio.org/terminal/ansipixels/tcolor"
	"fortio.org/tray/ray"
)

type stream struct {
	Idx        uint64      `json:"idx"`
	Seed       uint64      `json:"seed"`
	PCGUint64  []string    `json:"pcg_uint64"`  // math/rand/v2 PCG(idx, seed) raw outputs, decimal strings
	PCGNorm    []float64   `json:"pcg_norm"`    // mrand.New(NewPCG(idx, seed)).NormFloat64()
	Float64    []float64   `json:"float64"`     // rand.NewIdx(idx, seed).Float64()   (pins the seeding convention)
	UnitVector [][3]float64 `json:"unit_vector"` // rand.NewIdx(idx, seed).UnitVector()
	InDisc     [][2]float64 `json:"in_disc"`     // rand.NewIdx(idx, seed).InDisc(0.5)
	Vec3       [][3]float64 `json:"vec3"`        // rand.NewIdx(idx, seed).Vec3()
	Range      []float64   `json:"float64_range"` // rand.NewIdx(idx, seed).Float64Range(0.5, 1)
}

type out struct {
	GoVersion string             `json:"go_version"`
	GOARCH    string             `json:"goarch"`
	Streams   []stream           `json:"streams"`
	Srgb      map[string][]uint8 `json:"linear_to_srgb"` // "x": inputs as hex bit patterns -> outputs
	SrgbIn    []string           `json:"linear_to_srgb_inputs"`
	Exp       [][2]string        `json:"exp"` // (x, math.Exp(x)) hex bit patterns over the ziggurat wedge range
	Pow       [][2]string        `json:"pow"` // (x, math.Pow(x, 1/2.4))
	Tan       [][2]string        `json:"tan"` // (x, math.Tan(x)) for Camera.Initialize
	Objects   map[string]int     `json:"rich_scene_objects"` // seed -> len(RichScene(rand.New(seed)).Objects)
	Render    renderOut          `json:"render"`
}

type renderOut struct {
	Args   string  `json:"args"`
	Width  int     `json:"width"`
	Height int     `json:"height"`
	Pix    []uint8 `json:"pix"` // RGBA, benchmark -w 1 -seed 2 -width 40 -height 23 -r 4 -d 50
}

func bits(x float64) string { return "0x" + hex64(math.Float64bits(x)) }
func hex64(u uint64) string {
	const d = "0123456789abcdef"
	b := make([]byte, 16)
	for i := 15; i >= 0; i-- {
		b[i] = d[u&15]
		u >>= 4
	}
	return string(b)
}
func dec(u uint64) string {
	if u == 0 {
		return "0"
	}
	var b []byte
	for u > 0 {
		b = append([]byte{byte('0' + u%10)}, b...)
		u /= 10
	}
	return string(b)
}

func main() {
	o := out{GoVersion: runtime.Version(), GOARCH: runtime.GOARCH, Srgb: map[string][]uint8{}, Objects: map[string]int{}}
	const n = 64
	for _, seed := range []uint64{2, 7, 42} {
		for _, idx := range []uint64{0, 5, 1 << 40} {
			s := stream{Idx: idx, Seed: seed}
			p := mrand.NewPCG(idx, seed)
			for i := 0; i < n; i++ {
				s.PCGUint64 = append(s.PCGUint64, dec(p.Uint64()))
			}
			r := mrand.New(mrand.NewPCG(idx, seed))
			for i := 0; i < 4096; i++ {
				s.PCGNorm = append(s.PCGNorm, r.NormFloat64())
			}
			a := rand.NewIdx(int(idx), seed)
			for i := 0; i < n; i++ {
				s.Float64 = append(s.Float64, a.Float64())
			}
			b := rand.NewIdx(int(idx), seed)
			for i := 0; i < n; i++ {
				x, y, z := b.UnitVector()
				s.UnitVector = append(s.UnitVector, [3]float64{x, y, z})
			}
			c := rand.NewIdx(int(idx), seed)
			for i := 0; i < n; i++ {
				x, y := c.InDisc(0.5)
				s.InDisc = append(s.InDisc, [2]float64{x, y})
			}
			d := rand.NewIdx(int(idx), seed)
			for i := 0; i < n; i++ {
				x, y, z := d.Vec3()
				s.Vec3 = append(s.Vec3, [3]float64{x, y, z})
			}
			e := rand.NewIdx(int(idx), seed)
			for i := 0; i < n; i++ {
				s.Range = append(s.Range, e.Float64Range(0.5, 1))
			}
			o.Streams = append(o.Streams, s)
		}
	}
	// LinearToSrgb around every output threshold: bisect each k on the bit pattern, dump threshold -1/0/+1 ulp
	var outs []uint8
	for k := 1; k <= 255; k++ {
		lo, hi := uint64(0), math.Float64bits(1.0)
		for hi-lo > 1 {
			mid := lo + (hi-lo)/2
			if int(tcolor.LinearToSrgb(math.Float64frombits(mid))) >= k {
				hi = mid
			} else {
				lo = mid
			}
		}
		for _, u := range []uint64{hi - 1, hi, hi + 1} {
			o.SrgbIn = append(o.SrgbIn, "0x"+hex64(u))
			outs = append(outs, tcolor.LinearToSrgb(math.Float64frombits(u)))
		}
	}
	for _, x := range []float64{0, 1, 0.5, -0.5, 1.5, 0.25, 0.75, 0.0031308, math.Nextafter(0.0031308, 1)} {
		o.SrgbIn = append(o.SrgbIn, bits(x))
		outs = append(outs, tcolor.LinearToSrgb(x))
	}
	o.Srgb["out"] = outs
	for i := 0; i < 2000; i++ { // ziggurat wedge: exp(-x*x/2), x in (0, 3.45)
		x := -0.5 * (3.45 * float64(i) / 2000) * (3.45 * float64(i) / 2000)
		o.Exp = append(o.Exp, [2]string{bits(x), bits(math.Exp(x))})
	}
	for i := 1; i < 2000; i++ {
		x := float64(i) / 2000
		o.Pow = append(o.Pow, [2]string{bits(x), bits(math.Pow(x, 1/2.4))})
	}
	for _, deg := range []float64{10, 20, 30, 40, 45, 60, 75, 90, 120} {
		x := deg * math.Pi / 180 / 2
		o.Tan = append(o.Tan, [2]string{bits(x), bits(math.Tan(x))})
	}
	for seed := uint64(1); seed <= 12; seed++ {
		o.Objects[dec(seed)] = len(ray.RichScene(rand.New(seed)).Objects)
	}
	// the benchmark's own render, single worker (the -w 1 conformance semantics of DESIGN.md section 3)
	w, h := 40, 23
	rt := ray.New(w, h)
	rt.Camera = ray.RichSceneCamera()
	rt.MaxDepth, rt.NumRaysPerPixel, rt.NumWorkers, rt.Seed = 50, 4, 1, 2
	img := rt.Render(ray.RichScene(rand.New(2)))
	o.Render = renderOut{Args: "benchmark -w 1 -seed 2 -width 40 -height 23 -r 4 -d 50", Width: w, Height: h, Pix: img.Pix}
	enc := json.NewEncoder(os.Stdout)
	if err := enc.Encode(o); err != nil {
		panic(err)
	}
}
