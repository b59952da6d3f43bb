"""The ziggurat's wedge test compares with float32(exp(-x*x/2)) computed like Go (fp64 msun exp, then rounded). The device
decides most cases with a cheap fp32 exp first (pcg_norm in tray_device.cuh): accept if lhs < ef*0.999998, fall through to
the exact form if lhs < ef*1.000002, reject otherwise, with ef = expf((float)z). This checks that the band is wide enough
for ANY expf that is accurate to 3 ulp: the exact value always lies strictly inside (ef*0.999998, ef*1.000002)."""
import math

import numpy as np


def test_cheap_exp_band_always_contains_the_exact_value(O):
    L = O.lib()
    rs = np.random.RandomState(4)
    f32 = np.float32
    ulp = 2.0 ** -24
    worst = 0.0
    xs = np.concatenate([rs.uniform(0, 3.4426198, 20000), np.linspace(0, 3.4426198, 2001), -rs.uniform(0, 3.4426198, 2000)])
    for x in xs:
        z = -.5 * x * x
        exact = float(f32(L.oracle_go_exp(z)))                  # what the comparison is defined against
        t = math.exp(float(f32(z)))                             # true exp of the fp32-rounded argument
        for ef in (t * (1 - 3 * ulp), t, t * (1 + 3 * ulp)):    # any 3-ulp-accurate expf
            ef = float(f32(ef))
            lo, hi = float(f32(ef) * f32(0.999998)), float(f32(ef) * f32(1.000002))
            assert lo < exact < hi, (x, ef, exact)
            worst = max(worst, abs(exact / ef - 1))
    assert worst < 1e-6                                         # the band (2e-6) has a factor two in hand
