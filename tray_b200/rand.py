"""Host side of fortio.org/rand (third-party to the reference, go.mod:9): only what scene construction
needs -- New/NewIdx, Float64, Float64Range, Vec3. The per-sample generators run on the device."""
import secrets

_M64 = (1 << 64) - 1
_M128 = (1 << 128) - 1


class Rand:
    """rand.New(seed) / rand.NewIdx(idx, seed): Go math/rand/v2 PCG-DXSM seeded (uint64(idx), seed).
    Call sites: benchmark/benchmark.go:63, main.go:87, ray/tracer.go:121."""
    _MUL = (2549297995355413924 << 64) | 4865540595714422341
    _INC = (6364136223846793005 << 64) | 1442695040888963407

    def __init__(self, seed, idx=0):
        if seed == 0:  # "0 randomizes each time" (main.go:47, ray/tracer.go:32)
            seed = secrets.randbits(64) | 1
            idx = secrets.randbits(63)
        self.state = ((idx & _M64) << 64) | (seed & _M64)

    def Uint64(self):
        self.state = (self.state * self._MUL + self._INC) & _M128
        hi, lo = self.state >> 64, self.state & _M64
        hi ^= hi >> 32
        hi = (hi * 0xda942042e4dd58b5) & _M64
        hi ^= hi >> 48
        hi = (hi * (lo | 1)) & _M64
        return hi

    def Float64(self):
        return float(self.Uint64() & ((1 << 53) - 1)) / 9007199254740992.0

    def Float64Range(self, a, b):
        return a + (b - a) * self.Float64()

    def Vec3(self):
        return (self.Float64(), self.Float64(), self.Float64())


def New(seed):
    return Rand(seed, 0)


def NewIdx(idx, seed):
    return Rand(seed, idx)


