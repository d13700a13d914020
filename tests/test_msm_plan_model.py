"""CPU model of the index arithmetic of the GPU's Pippenger (pcd_b200/csrc/msm.cuh, wec.cuh), over the additive group Z_q
instead of a curve: signed digits on balanced windows (msm_win_start / msm_digits_kernel), ONE bucket set shared by all
windows of a precomputed table T[w][i] = 2^(start_w) P_i, unit buckets for scalars equal to 1, bucket parts, and the
bucket reduction by groups of L buckets (running sums + [t L] * plain sum; unit groups with weight 1: wec_reduce_kernel).
The result must equal sum_i k_i P_i -- what `VariableBaseMSM::multi_scalar_mul` computes (ark-ec, reached from
/root/reference/src/ec_cycle_pcd/mod.rs:171,179 through `Groth16::prove`).  It pins the window starts, the carry into
the top window, the unit-bucket digit code and the reduction weights; the kernels themselves are checked on the GPU
(tests/test_gpu_parity.py, tests/test_gpu_baseline_sizes.py)."""
import random

import pytest

SCALAR_BITS = 298          # msm.cuh: MSM_SCALAR_BITS (both scalar fields of the cycle are 298-bit)
Q = (1 << 127) - 1         # the toy group Z_Q
R4 = 475922286169261325753349249653048451545124878552823515553267735739164647307408490559963137


def num_windows(c):        # api.cu: msm_num_windows_c
    return (SCALAR_BITS + 1 + c - 1) // c


def win_start(j, c, nwin, balanced):   # msm.cuh: msm_win_start
    if not balanced:
        return j * c
    base, rem = divmod(SCALAR_BITS + 1, nwin)
    return j * base + min(j, rem)


def digits(k, c, nwin, shared):        # msm.cuh: msm_digits_kernel (the signed recoding loop)
    out, carry = [], 0
    for w in range(nwin):
        bit = win_start(w, c, nwin, shared)
        nxt = win_start(w + 1, c, nwin, shared) if w + 1 < nwin else (SCALAR_BITS + 1 if shared else bit + c)
        cw = nxt - bit
        d = ((k >> bit) & ((1 << cw) - 1)) + carry
        if d > (1 << (cw - 1)):
            out.append(d - (1 << cw))
            carry = 1
        else:
            out.append(d)
            carry = 0
    assert carry == 0, "a 298-bit scalar never carries out of the top window"
    return out


def msm_model(points, scalars, c, unit_k, log_l, split):
    nwin = num_windows(c)
    B = 1 << (c - 1)
    starts = [win_start(w, c, nwin, True) for w in range(nwin)]
    table = [[(p << s) % Q for p in points] for s in starts]      # precompute_kernel: T[w][i] = 2^start_w P_i
    entries = [[] for _ in range(B + unit_k)]                      # msm_scatter_kernel: (table index, sign) per bucket
    for i, k in enumerate(scalars):
        if unit_k and k == 1:                                      # unit bucket: digit code B + 1 + i mod unit_k
            entries[B + (i & (unit_k - 1))].append((0, i, 1))
            continue
        for w, d in enumerate(digits(k, c, nwin, True)):
            if d:
                entries[abs(d) - 1].append((w, i, 1 if d > 0 else -1))
    buckets = []
    for ent in entries:                                            # msm_accumulate_kernel + msm_fold_parts_kernel
        parts = [0] * split
        for part in range(split):
            lo, hi = len(ent) * part // split, len(ent) * (part + 1) // split
            for w, i, sgn in ent[lo:hi]:
                parts[part] = (parts[part] + sgn * table[w][i]) % Q
        buckets.append(sum(parts) % Q)
    L = 1 << log_l
    assert B % L == 0 and unit_k % L == 0
    total = 0
    for t in range(B // L):                                        # wec_reduce_kernel: weighted groups
        run = acc = 0
        for j in range(L - 1, -1, -1):
            run = (run + buckets[t * L + j]) % Q
            acc = (acc + run) % Q
        total = (total + acc + (t * L) * run) % Q
    for t in range(unit_k // L):                                   # ... and unit groups, weight 1
        total = (total + sum(buckets[B + t * L:B + (t + 1) * L])) % Q
    return total


@pytest.mark.parametrize("c,unit_k,log_l,split", [(6, 0, 1, 1), (7, 0, 3, 4), (10, 64, 3, 4), (12, 64, 2, 8), (15, 256, 3, 4)])
def test_model_equals_direct_sum(c, unit_k, log_l, split):
    rnd = random.Random(1000 * c + unit_k)
    n = 300
    points = [rnd.randrange(1, Q) for _ in range(n)]
    edge = [0, 1, 2, R4 - 1, R4 - 2, 1 << 297, (1 << 298) - 1 if (1 << 298) - 1 < R4 else R4 - 3, (1 << (c - 1)), (1 << (c - 1)) + 1]
    scalars = [rnd.choice(edge) if rnd.random() < 0.3 else (rnd.choice([0, 1]) if rnd.random() < 0.4 else rnd.randrange(R4))
               for _ in range(n)]
    expect = sum(k * p for k, p in zip(scalars, points)) % Q
    assert msm_model(points, scalars, c, unit_k, log_l, split) == expect


def test_window_starts_cover_the_scalar():
    for c in range(4, 22):
        nwin = num_windows(c)
        starts = [win_start(w, c, nwin, True) for w in range(nwin)] + [SCALAR_BITS + 1]
        widths = [b - a for a, b in zip(starts, starts[1:])]
        assert starts[0] == 0 and all(0 < w <= c for w in widths) and max(widths) - min(widths) <= 1, (c, widths)
