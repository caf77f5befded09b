"""What does curandSetGeneratorOffset(n) select for XORWOW + GenerateUniformDouble?  (skip_curand, src/ising3d_gpu_m.f90:72-77)
Run on the GPU box: python tools/curand_probe.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from test_gpu_curand_stream import Xorwow  # noqa: E402

n = 1 << 16
big = Xorwow(42).generate(8 * n)
two = Xorwow(42)
f, s = two.generate(n), two.generate(n)
print("call1 == big[:n]", np.array_equal(f, big[:n]), " call2 == big[n:2n]", np.array_equal(s, big[n:2 * n]))
for off in (1, 2, 4096, n // 2, n, 2 * n):
    g = Xorwow(42)
    g.set_offset(off)
    b = g.generate(n)
    hit = [k for k in range(0, 7 * n + 1) if b[0] == big[k]][:3]
    ok = [k for k in hit if np.array_equal(b, big[k:k + n])]
    print(f"offset {off}: first value found at big index {hit}, whole block equal at {ok}")
m = 1000   # a call size that is not a multiple of anything
t = Xorwow(42)
c1, c2 = t.generate(m), t.generate(m)
print("odd sizes: call1 == big[:m]", np.array_equal(c1, big[:m]), " call2 == big[m:2m]", np.array_equal(c2, big[m:2 * m]))
