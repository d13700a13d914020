import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import c_oracle as co, codec, synth, pcd_b200
ctx = pcd_b200.Context(0)
for curve in (1, 2, 3):
    n = 64
    k = synth.random_scalars(n, curve, 44)
    k[0] = 0
    k[1] = codec.int_to_limbs(1)
    k[2] = codec.int_to_limbs(2)
    k[3] = codec.int_to_limbs(15)
    k[4] = codec.int_to_limbs(16)
    k[5] = codec.int_to_limbs(17)
    k[6] = codec.int_to_limbs(1 << 64)
    k[7] = codec.int_to_limbs((1 << 200) + 5)
    g = synth.generator_limbs(curve)
    a = ctx.fixed_base_mul(curve, g, k)
    b = co.fixed_base_mul(curve, g, k)
    bad = [i for i in range(n) if not np.array_equal(a[i], b[i])]
    print("curve", curve, "bad rows", bad[:20], len(bad))
