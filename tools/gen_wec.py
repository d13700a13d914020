#!/usr/bin/env python3
"""Generator of pcd_b200/csrc/wec_programs.cuh: the lane schedules of the warp-cooperative group law.

The latency-bound tails of a proof (bucket reduction, s*g_a + r*g1_b, window combination, normalisation to affine)
are chains of dependent group operations on a handful of points; run by one thread each, an operation is 10..82
dependent 298-bit Montgomery products (0.7 us apiece on B200).  The products INSIDE one group operation are mostly
independent of each other, so here a group of G lanes of a warp computes ONE operation: the formula (the same XYZZ
formulas as ec.cuh, over Fq / Fq2 / Fq3 with the same Karatsuba / Chung-Hasan products as fpx.cuh) is traced into a DAG
of base-field operations and list-scheduled into steps; in a step every lane executes at most one base-field
instruction on operands in shared memory (pcd_b200/csrc/wec.cuh interprets the schedule).  A group addition then costs
its multiplicative DEPTH (4..5 products) instead of its product COUNT.

Instruction set (one instruction per lane and ROW, the group synchronises after every row; all the instructions of a
row have the same opcode, so the lanes of a warp never diverge inside a row):
  MUL  d = a * b     Montgomery product; operands may be UNREDUCED (< 2^10 p each), the result is canonical (< p)
  LIN  d = a + k b   k a small signed constant; NO reduction: the generator tracks for every value a bound m (value
                     < m p) and turns a negative k into + (2^j p - |k| b) with 2^j >= |k| m_b, so nothing ever
                     underflows or needs a conditional subtraction; a may be the constant 0
  RED  d = a mod p   canonical representative of a value < 2^m p (m conditional subtractions): only where a value
                     leaves the program (outputs) or would exceed the product's operand range
  INV  d = 1 / a     (binary extended Euclid, fp.cuh)
Additions, subtractions, doublings, negations and multiplications by the small curve / non-residue constants of the
formulas all lower to LIN, and a `LIN(x, k1 * y)` whose y is itself `k2 * u` is fused into `LIN(x, (k1 k2) u)`.
Instruction = two u32 words: w0 = op << 28 | dst << 18 | a << 9 | b;  w1 = (k & 0xff) | j << 8 | m << 13 | zero_a << 18.
Slot: 0..15 = X (the point updated in place), 16..31 = Y (the point added), 32.. = T (temporaries; the first ones
carry values between the two halves of an addition).  One slot = one base-field element (ten u32 words, Montgomery
form).

  python tools/gen_wec.py            # rewrite pcd_b200/csrc/wec_programs.cuh
  python tools/gen_wec.py --check    # exit 1 if the committed header differs from what this script generates

tests/test_wec_programs.py interprets the generated schedules with Python integers and compares every program with
the oracle's group law, so the schedules are checked without a GPU; tests/test_gpu_wec.py runs the interpreter on B200.
"""
from __future__ import annotations

import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "pcd_b200", "csrc", "wec_programs.cuh")

OP_NOP, OP_MUL, OP_LIN, OP_RED, OP_INV, OP_CPY = range(6)
OP_NAMES = ["nop", "mul", "lin", "red", "inv", "cpy"]
MUL_BOUND = 1 << 20   # product of the operands' bounds a Montgomery product accepts (T < p (1 + 2^-2) < 2 p)
MAX_BOUND = 1 << 21   # a value must stay below 2^21 p < 2^320
REG_X, REG_Y, REG_T = 0, 1, 2
MAX_T = 224
X_BASE, Y_BASE, T_BASE = 0, 16, 32


# ---- tracing ---------------------------------------------------------------------------------------------------
class Node:
    __slots__ = ("op", "a", "b", "k", "id", "pin", "name")

    def __init__(self, g, op, a=None, b=None, k=0, pin=None, name=""):
        self.op, self.a, self.b, self.k, self.pin, self.name = op, a, b, k, pin, name
        self.id = len(g.nodes)
        g.nodes.append(self)


class Graph:
    def __init__(self):
        self.nodes = []
        self.cse = {}

    def inp(self, region, index, name=""):
        return Node(self, "in", pin=(region, index), name=name)

    def op(self, op, a, b=None, k=0):
        if op in ("mul", "add") and b is not None and b.id < a.id:
            a, b = b, a
        key = (op, a.id, b.id if b is not None else -1, k)
        if key not in self.cse:
            self.cse[key] = Node(self, op, a, b, k)
        return self.cse[key]


class B:
    """traced base-field element"""

    def __init__(self, g, n):
        self.g, self.n = g, n

    def __mul__(self, o):
        return B(self.g, self.g.op("mul", self.n, o.n))

    def __add__(self, o):
        return B(self.g, self.g.op("add", self.n, o.n))

    def __sub__(self, o):
        return B(self.g, self.g.op("sub", self.n, o.n))

    def sqr(self):
        return self * self

    def subd(self, o):  # self - 2 o
        return B(self.g, self.g.op("subd", self.n, o.n))

    def dbl(self):
        return B(self.g, self.g.op("dbl", self.n))

    def neg(self):
        return B(self.g, self.g.op("neg", self.n))

    def mulk(self, k):
        if k == 1:
            return self
        if k == 2:
            return self.dbl()
        return B(self.g, self.g.op("mulk", self.n, None, k))

    def inv(self):
        return B(self.g, self.g.op("inv", self.n))


class E:
    """traced extension-field element: tuple of base elements; formulas follow pcd_b200/csrc/fpx.cuh"""

    def __init__(self, c, nr):
        self.c, self.nr = tuple(c), nr

    @property
    def k(self):
        return len(self.c)

    def _lin(self, o, f):
        return E([f(x, y) for x, y in zip(self.c, o.c)], self.nr)

    def __add__(self, o):
        return self._lin(o, lambda x, y: x + y)

    def __sub__(self, o):
        return self._lin(o, lambda x, y: x - y)

    def subd(self, o):
        return self._lin(o, lambda x, y: x.subd(y))

    def dbl(self):
        return E([x.dbl() for x in self.c], self.nr)

    def mulk(self, k):
        return E([x.mulk(k) for x in self.c], self.nr)

    def __mul__(self, o):
        a, b, nr = self.c, o.c, self.nr
        if self.k == 1:
            return E([a[0] * b[0]], nr)
        if self.k == 2:  # Karatsuba, 3 products
            v0, v1 = a[0] * b[0], a[1] * b[1]
            c1 = (a[0] + a[1]) * (b[0] + b[1]) - v0 - v1
            return E([v0 + v1.mulk(nr), c1], nr)
        v0, v1, v2 = a[0] * b[0], a[1] * b[1], a[2] * b[2]  # Karatsuba, 6 products
        w12, w01, w02 = (a[1] + a[2]) * (b[1] + b[2]), (a[0] + a[1]) * (b[0] + b[1]), (a[0] + a[2]) * (b[0] + b[2])
        # the same sums as fpx.cuh, associated as balanced trees (two dependent linear rows instead of four):
        #   c0 = v0 + nr (w12 - v1 - v2) = (v0 - nr v2) + nr (w12 - v1)
        #   c1 = w01 - v0 - v1 + nr v2   = (w01 - v0) - (v1 - nr v2)
        #   c2 = w02 - v0 - v2 + v1      = (w02 - v0) + (v1 - v2)
        c0 = (v0 - v2.mulk(nr)) + (w12 - v1).mulk(nr)
        c1 = (w01 - v0) - (v1 - v2.mulk(nr))
        c2 = (w02 - v0) + (v1 - v2)
        return E([c0, c1, c2], nr)

    def sqr(self):
        a, nr = self.c, self.nr
        if self.k == 1:
            return E([a[0] * a[0]], nr)
        if self.k == 2:  # complex squaring, 2 products
            ab = a[0] * a[1]
            t = (a[0] + a[1]) * (a[0] + a[1].mulk(nr))
            return E([t - ab - ab.mulk(nr), ab.dbl()], nr)
        s0 = a[0] * a[0]  # Chung-Hasan SQR2, 5 products
        s1 = (a[0] * a[1]).dbl()
        t2 = a[0] - a[1] + a[2]
        s2 = t2 * t2
        s3 = (a[1] * a[2]).dbl()
        s4 = a[2] * a[2]
        return E([s0 + s3.mulk(nr), s1 + s4.mulk(nr), (s1 + s2) + (s3 - (s0 + s4))], nr)

    def inv(self):
        a, nr = self.c, self.nr
        if self.k == 1:
            return E([a[0].inv()], nr)
        if self.k == 2:
            n = a[0] * a[0] - (a[1] * a[1]).mulk(nr)
            ni = n.inv()
            return E([a[0] * ni, (a[1] * ni).neg()], nr)
        t0 = a[0] * a[0] - (a[1] * a[2]).mulk(nr)
        t1 = (a[2] * a[2]).mulk(nr) - a[0] * a[1]
        t2 = a[1] * a[1] - a[0] * a[2]
        n = a[0] * t0 + (a[2] * t1 + a[1] * t2).mulk(nr)
        ni = n.inv()
        return E([t0 * ni, t1 * ni, t2 * ni], nr)


# curve id -> (extension degree, non-residue, mul_a, default group size)
def _mul_a_0(v):
    return v.dbl()


def _mul_a_1(v):
    return v.mulk(34)


def _mul_a_2(v):
    return v.mulk(11)


def _mul_a_3(v):
    return E([v.c[1].mulk(55), v.c[2].mulk(55), v.c[0].mulk(11)], v.nr)


CURVES = {
    0: dict(name="MNT4_G1", k=1, nr=0, mul_a=_mul_a_0, G=4),
    1: dict(name="MNT4_G2", k=2, nr=17, mul_a=_mul_a_1, G=16),
    2: dict(name="MNT6_G1", k=1, nr=0, mul_a=_mul_a_2, G=4),
    3: dict(name="MNT6_G2", k=3, nr=5, mul_a=_mul_a_3, G=32),
}


def _ext_in(g, region, first, k, nr, name):
    return E([B(g, g.inp(region, first + i, "%s%d" % (name, i))) for i in range(k)], nr)


def trace(curve, prog):
    """returns (graph, outputs): outputs = list of (node, (region, index))"""
    cv = CURVES[curve]
    k, nr = cv["k"], cv["nr"]
    g = Graph()
    outs = []

    def out(e, region, first):
        for i, c in enumerate(e.c):
            outs.append((c.n, (region, first + i)))

    X = lambda j, nm: _ext_in(g, REG_X, j * k, k, nr, nm)
    Y = lambda j, nm: _ext_in(g, REG_Y, j * k, k, nr, nm)
    T = lambda j, nm: _ext_in(g, REG_T, j * k, k, nr, nm)
    if prog == "add1":  # T <- P, R, U1, S1
        x1, y1, zz1, zzz1 = X(0, "x"), X(1, "y"), X(2, "zz"), X(3, "zzz")
        x2, y2, zz2, zzz2 = Y(0, "x"), Y(1, "y"), Y(2, "zz"), Y(3, "zzz")
        U1, U2, S1, S2 = x1 * zz2, x2 * zz1, y1 * zzz2, y2 * zzz1
        out(U2 - U1, REG_T, 0)
        out(S2 - S1, REG_T, k)
        out(U1, REG_T, 2 * k)
        out(S1, REG_T, 3 * k)
    elif prog == "add2":  # X <- the sum, from add1's T values
        P, R, U1, S1 = (T(j, n) for j, n in enumerate(("P", "R", "U1", "S1")))
        ZZa, ZZZa = X(2, "zz") * Y(2, "zz"), X(3, "zzz") * Y(3, "zzz")
        PP = P.sqr()
        PPP = P * PP
        Q = U1 * PP
        x3 = (R.sqr() - PPP).subd(Q)
        out(x3, REG_X, 0)
        out(R * (Q - x3) - S1 * PPP, REG_X, k)
        out(ZZa * PP, REG_X, 2 * k)
        out(ZZZa * PPP, REG_X, 3 * k)
    elif prog == "madd1":  # Y affine: T <- P, R
        x1, y1, zz1, zzz1 = X(0, "x"), X(1, "y"), X(2, "zz"), X(3, "zzz")
        x2, y2 = Y(0, "x"), Y(1, "y")
        out(x2 * zz1 - x1, REG_T, 0)
        out(y2 * zzz1 - y1, REG_T, k)
    elif prog == "madd2":
        x1, y1, zz1, zzz1 = X(0, "x"), X(1, "y"), X(2, "zz"), X(3, "zzz")
        P, R = T(0, "P"), T(1, "R")
        PP = P.sqr()
        PPP = P * PP
        Q = x1 * PP
        x3 = (R.sqr() - PPP).subd(Q)
        out(x3, REG_X, 0)
        out(R * (Q - x3) - y1 * PPP, REG_X, k)
        out(zz1 * PP, REG_X, 2 * k)
        out(zzz1 * PPP, REG_X, 3 * k)
    elif prog == "dbl":
        x, y, zz, zzz = X(0, "x"), X(1, "y"), X(2, "zz"), X(3, "zzz")
        U = y.dbl()
        V = U.sqr()
        W = U * V
        S = x * V
        xx = x.sqr()
        M = xx.mulk(3) + cv["mul_a"](zz.sqr())
        x3 = M.sqr().subd(S)
        out(x3, REG_X, 0)
        out(M * (S - x3) - W * y, REG_X, k)
        out(V * zz, REG_X, 2 * k)
        out(W * zzz, REG_X, 3 * k)
    elif prog == "toaff":  # X[0..2k) <- affine (x, y)
        x, y, zz, zzz = X(0, "x"), X(1, "y"), X(2, "zz"), X(3, "zzz")
        i = zzz.inv()
        t = zz * i
        out(x * t.sqr(), REG_X, 0)
        out(y * i, REG_X, k)
    else:
        raise ValueError(prog)
    return g, outs


PROGRAMS = ["add1", "add2", "madd1", "madd2", "dbl", "toaff"]


# ---- lowering: formulas -> MUL / LIN / RED / INV with bounds -------------------------------------------------------
class LNode:
    __slots__ = ("op", "a", "b", "k", "j", "m", "bound", "id", "pin")

    def __init__(self, lg, op, a=None, b=None, k=0, pin=None):
        self.op, self.a, self.b, self.k, self.pin = op, a, b, k, pin
        self.j = self.m = 0
        self.bound = 1
        self.id = len(lg)
        lg.append(self)


def lower(g, outs):
    """traced DAG -> (list of LNode, outputs); only nodes an output depends on are lowered"""
    needed = set()
    stack = [n for n, _ in outs]
    while stack:
        n = stack.pop()
        if n.id in needed:
            continue
        needed.add(n.id)
        for o in (n.a, n.b):
            if o is not None:
                stack.append(o)
    uses = {}
    for i in needed:
        n = g.nodes[i]
        for o in (n.a, n.b):
            if o is not None:
                uses[o.id] = uses.get(o.id, 0) + 1
    for n, _ in outs:
        uses[n.id] = uses.get(n.id, 0) + 1
    lg = []
    m = {}  # traced node id -> LNode
    scaled = {}  # LNode id -> (source LNode, k): the node is k * source with a == zero

    def lin(a, b, k):
        """a + k b with b possibly itself a pure scaling used once"""
        if b.id in scaled and scaled[b.id][2] == 1 and abs(k * scaled[b.id][1]) <= 127:
            src, k2, _ = scaled[b.id]
            b, k = src, k * k2
            # the fused scaling node stays in lg but is no longer needed: pruned by the scheduler's reachability
        n = LNode(lg, "lin", a, b, k)
        return n

    for i in sorted(needed):
        t = g.nodes[i]
        if t.op == "in":
            m[i] = LNode(lg, "in", pin=t.pin)
        elif t.op == "mul":
            m[i] = LNode(lg, "mul", m[t.a.id], m[t.b.id])
        elif t.op == "inv":
            m[i] = LNode(lg, "inv", m[t.a.id])
        elif t.op == "add":
            a, b = m[t.a.id], m[t.b.id]
            if a.id in scaled and scaled[a.id][2] == 1 and b.id not in scaled:
                a, b = b, a  # put the scaling on the b side so that it fuses
            m[i] = lin(a, b, 1)
        elif t.op == "sub":
            m[i] = lin(m[t.a.id], m[t.b.id], -1)
        elif t.op == "subd":
            m[i] = lin(m[t.a.id], m[t.b.id], -2)
        elif t.op in ("dbl", "neg", "mulk"):
            k = {"dbl": 2, "neg": -1}.get(t.op, t.k)
            src = m[t.a.id]
            if src.id in scaled and scaled[src.id][2] == 1 and abs(k * scaled[src.id][1]) <= 127:
                k, src = k * scaled[src.id][1], scaled[src.id][0]
            n = LNode(lg, "lin", None, src, k)
            scaled[n.id] = (src, k, uses.get(i, 0))
            m[i] = n
        else:
            raise ValueError(t.op)
    louts = [(m[n.id], slot) for n, slot in outs]
    return lg, louts


def finalize_bounds(lg, louts):
    """bounds in topological order; RED nodes where a product's operands are too large and on every output that is not
    canonical already.  Returns (lg, louts) with REDs appended / operands rewired."""
    red_of = {}

    def red(n):
        if n.bound <= 1:
            return n
        if n.id not in red_of:
            r = LNode(lg, "red", n)
            r.m = max(1, (n.bound - 1).bit_length())
            r.bound = 1
            red_of[n.id] = r
        return red_of[n.id]

    for n in list(lg):
        if n.op == "in":
            n.bound = 1
        elif n.op == "mul":
            while n.a.bound * n.b.bound > MUL_BOUND:
                if n.a.bound >= n.b.bound:
                    n.a = red(n.a)
                else:
                    n.b = red(n.b)
            n.bound = 1
        elif n.op == "inv":
            n.a = red(n.a)
            n.bound = 1
        elif n.op == "lin":
            kb = abs(n.k) * n.b.bound
            ab = n.a.bound if n.a is not None else 0
            if n.k < 0:
                n.j = max(0, (kb - 1).bit_length())  # 2^j >= |k| m_b
                n.bound = ab + (1 << n.j)
            else:
                n.bound = ab + kb
            assert n.bound < MAX_BOUND and n.j < 32, "bound overflow"
    louts = [(red(n), slot) for n, slot in louts]
    # REDs were appended after their sources but possibly after their users in id order: re-sort topologically
    order, seen = [], set()

    def visit(n):
        if n.id in seen:
            return
        seen.add(n.id)
        for o in (n.a, n.b):
            if o is not None:
                visit(o)
        order.append(n)

    for n, _ in louts:
        visit(n)
    for k, n in enumerate(order):
        n.id = k
    return order, louts


# ---- scheduling ------------------------------------------------------------------------------------------------
KIND = {"mul": "mul", "inv": "inv", "lin": "lin", "red": "red"}


def schedule(lg, louts, G):
    """List scheduling by dependency level.  Every row holds instructions of ONE opcode (no divergence inside a row):
    ready LIN instructions first, then RED, then products (up to G, deepest remaining chain first), an inversion alone."""
    users = {n.id: [] for n in lg}
    for n in lg:
        for o in (n.a, n.b):
            if o is not None:
                users[o.id].append(n.id)
    by_id = {n.id: n for n in lg}
    height = {}

    def h(i):
        if i not in height:
            n = by_id[i]
            w = 10 if n.op in ("mul", "inv") else 1
            height[i] = w + max([h(u) for u in users[i]], default=0)
        return height[i]

    done = {n.id for n in lg if n.op == "in"}
    todo = [n.id for n in lg if n.op != "in"]
    ready = lambda i: all(o is None or o.id in done for o in (by_id[i].a, by_id[i].b))
    rows = []
    while todo:
        pick = None
        # reductions that feed a product go as soon as they are ready; the outputs' reductions wait until nothing
        # else is, so that they share one row
        for kind, want_users in (("lin", None), ("red", True), ("mul", None), ("inv", None), ("red", False)):
            cand = sorted((i for i in todo if by_id[i].op == kind and ready(i) and
                           (want_users is None or bool(users[i]) == want_users)), key=lambda i: -h(i))
            if cand:
                pick = cand[:1] if kind == "inv" else cand[:G]
                rows.append((kind, pick))
                break
        assert pick, "scheduler stuck"
        done.update(pick)
        todo = [i for i in todo if i not in done]
    return rows, by_id


def allocate(lg, louts, rows, by_id, G):
    """Slots: inputs are pinned; an output goes straight to its slot when the value living there is dead, else to a
    temporary with a copy row at the end; a temporary is reused after the ROW of its last use (a slot is never read and
    written by different lanes inside one row)."""
    last_use = {}
    for s, (_, ids) in enumerate(rows):
        for i in ids:
            n = by_id[i]
            for o in (n.a, n.b):
                if o is not None:
                    last_use[o.id] = s
    out_of = {}
    for n, slot in louts:
        out_of.setdefault(n.id, []).append(slot)
    out_slots = {slot for _, slot in louts}
    pinned_until = {}
    for n in lg:
        if n.op == "in":
            pinned_until[n.pin] = max(pinned_until.get(n.pin, -1), last_use.get(n.id, -1))
    loc = {n.id: n.pin for n in lg if n.op == "in"}
    t_pinned = {n.pin[1] for n in lg if n.op == "in" and n.pin[0] == REG_T}
    t_pinned |= {s[1] for s in out_slots if s[0] == REG_T}
    free_t = [i for i in range(MAX_T) if i not in t_pinned]
    release = {}
    final_copies = []
    max_t = max(t_pinned, default=-1)
    out_rows = []
    for s, (kind, ids) in enumerate(rows):
        row = []
        for i in ids:
            n = by_id[i]
            dst = None
            for slot in out_of.get(i, []):
                if pinned_until.get(slot, -1) < s and dst is None:
                    dst = slot
                    pinned_until[slot] = 10 ** 9
            if dst is None:
                assert free_t, "out of temporaries"
                t = free_t.pop(0)
                max_t = max(max_t, t)
                dst = (REG_T, t)
                if i not in out_of:
                    release.setdefault(last_use.get(i, s), []).append(t)
            loc[i] = dst
            for slot in out_of.get(i, []):
                if slot != dst:
                    final_copies.append((slot, i))
            row.append(dict(op=kind, d=dst, a=loc[n.a.id] if n.a is not None else None,
                            b=loc[n.b.id] if n.b is not None else None, k=n.k, j=n.j, m=n.m))
        out_rows.append((kind, row))
        for t in release.pop(s, []):
            free_t.append(t)
        free_t.sort()
    for n, slot in louts:
        if loc[n.id] != slot and (slot, n.id) not in final_copies:
            final_copies.append((slot, n.id))
    if final_copies:
        srcs = {loc[i] for _, i in final_copies}
        dsts = [sl for sl, _ in final_copies]
        assert not (srcs & set(dsts)), "final copies alias"
        for c in range(0, len(final_copies), G):
            out_rows.append(("cpy", [dict(op="cpy", d=slot, a=loc[i], b=None, k=0, j=0, m=0)
                                     for slot, i in final_copies[c:c + G]]))
    return out_rows, max_t + 1


def enc_slot(s):
    if s is None:
        return 0
    base = {REG_X: X_BASE, REG_Y: Y_BASE, REG_T: T_BASE}[s[0]]
    lim = {REG_X: 16, REG_Y: 16, REG_T: MAX_T}[s[0]]
    assert 0 <= s[1] < lim
    return base + s[1]


def encode(rows, G):
    """two u32 per lane and row: [row][lane][2]"""
    words = []
    opc = {n: i for i, n in enumerate(OP_NAMES)}
    for kind, row in rows:
        assert len(row) <= G
        for ins in row:
            w0 = (opc[ins["op"]] << 28) | (enc_slot(ins["d"]) << 18) | (enc_slot(ins["a"]) << 9) | enc_slot(ins["b"])
            assert -128 <= ins["k"] <= 127 and 0 <= ins["j"] < 32 and 0 <= ins["m"] < 32
            zero_a = 1 if (ins["op"] == "lin" and ins["a"] is None) else 0
            w1 = (ins["k"] & 0xff) | (ins["j"] << 8) | (ins["m"] << 13) | (zero_a << 18)
            words += [w0, w1]
        words += [0, 0] * (G - len(row))
    return words


def build(curve, prog, G=None):
    G = G or CURVES[curve]["G"]
    g, outs = trace(curve, prog)
    lg, louts = lower(g, outs)
    lg, louts = finalize_bounds(lg, louts)
    rows, by_id = schedule(lg, louts, G)
    arows, nt = allocate(lg, louts, rows, by_id, G)
    count = lambda k: sum(1 for kind, _ in arows if kind == k)
    return dict(curve=curve, prog=prog, G=G, rows=arows, words=encode(arows, G), nrows=len(arows), ntemps=nt,
                mul_rows=count("mul") + count("inv"), lin_rows=count("lin") + count("cpy"), red_rows=count("red"),
                max_bound=max(n.bound for n in lg))


# ---- Python interpreter (tests) ----------------------------------------------------------------------------------
def interpret(built, p, X, Y, T=None):
    """Run a schedule on Python integers with the EXACT semantics of the GPU interpreter: LIN does not reduce, RED
    subtracts 2^s p conditionally for s = m-1 .. 0, MUL takes unreduced operands and returns the canonical product.
    Checks the invariants the GPU relies on: no LIN underflows, every value stays below 2^320, a product's operands
    are within its range, RED's input is below 2^m p, and inside a row no slot written by one lane is read or written
    by another.  X, Y, T: lists of ints (modified in place)."""
    G = built["G"]
    T = T if T is not None else [0] * MAX_T
    X += [0] * (16 - len(X))
    Y += [0] * (16 - len(Y))
    mem = X + Y + T
    words = built["words"]
    for r in range(built["nrows"]):
        reads = [set() for _ in range(G)]
        writes = [set() for _ in range(G)]
        pending = []
        ops = set()
        for lane in range(G):
            w0, w1 = words[2 * (r * G + lane)], words[2 * (r * G + lane) + 1]
            op, d, a, b = w0 >> 28, (w0 >> 18) & 511, (w0 >> 9) & 511, w0 & 511
            if op == OP_NOP:
                continue
            ops.add(op)
            k = w1 & 0xff
            k = k - 256 if k >= 128 else k
            j, m, zero_a = (w1 >> 8) & 31, (w1 >> 13) & 31, (w1 >> 18) & 1
            if op == OP_MUL:
                av, bv = mem[a], mem[b]
                reads[lane] |= {a, b}
                assert av * bv < (MUL_BOUND << 2) * p * p, "product operands out of range"
                res = av * bv % p
            elif op == OP_LIN:
                bv = mem[b]
                reads[lane].add(b)
                t = abs(k) * bv
                if k < 0:
                    t = (p << j) - t
                    assert t >= 0, "LIN underflow"
                if not zero_a:
                    t += mem[a]
                    reads[lane].add(a)
                assert t < (1 << 320), "LIN overflow"
                res = t
            elif op == OP_RED:
                v = mem[a]
                reads[lane].add(a)
                assert v < (p << m), "RED input too large"
                for sft in range(m - 1, -1, -1):
                    if v >= (p << sft):
                        v -= p << sft
                assert v < p
                res = v
            elif op == OP_INV:
                reads[lane].add(a)
                assert mem[a] < p
                res = pow(mem[a], p - 2, p)
            elif op == OP_CPY:
                reads[lane].add(a)
                res = mem[a]
            else:
                raise ValueError(op)
            pending.append((d, res))
            writes[lane].add(d)
        assert len(ops) <= 1, "mixed opcodes in a row"
        for l1 in range(G):
            for l2 in range(G):
                if l1 != l2:
                    assert not (writes[l1] & (reads[l2] | writes[l2])), "cross-lane hazard in a row"
        for d, res in pending:
            mem[d] = res
    X[:] = mem[0:16]
    Y[:] = mem[16:32]
    T[:] = mem[32:]
    return X, Y, T


# ---- header ------------------------------------------------------------------------------------------------------
def header():
    lines = ["// GENERATED by tools/gen_wec.py -- do not edit (python tools/gen_wec.py --check verifies it is current).",
             "// Lane schedules of the warp-cooperative group law (wec.cuh): two u32 per lane and row,",
             "// w0 = op << 28 | dst << 18 | a << 9 | b;  w1 = (k & 0xff) | j << 8 | m << 13 | zero_a << 18;",
             "// slots 0..15 X, 16..31 Y, 32.. T; ops: " + ", ".join("%d %s" % (i, n) for i, n in enumerate(OP_NAMES)) + ".",
             "#pragma once", "#include \"prims.cuh\"", ""]
    summary = []
    for curve in sorted(CURVES):
        cv = CURVES[curve]
        nt = 0
        for prog in PROGRAMS:
            b = build(curve, prog)
            nt = max(nt, b["ntemps"])
            name = "WEC_%s_%s" % (cv["name"], prog.upper())
            lines.append("// %s %s: %d rows (%d product, %d linear, %d reduction), %d temporaries, %d lanes, largest bound %d p" %
                         (cv["name"], prog, b["nrows"], b["mul_rows"], b["lin_rows"], b["red_rows"], b["ntemps"], b["G"],
                          b["max_bound"]))
            lines.append("__device__ const u32 %s[%d] = {" % (name, len(b["words"])))
            ws = b["words"]
            for c in range(0, len(ws), 8):
                lines.append("  " + ", ".join("0x%08xu" % w for w in ws[c:c + 8]) + ",")
            lines.append("};")
            lines.append("static constexpr int %s_ROWS = %d;" % (name, b["nrows"]))
            summary.append((cv["name"], prog, b["nrows"], b["mul_rows"], b["lin_rows"], b["red_rows"], b["ntemps"], b["max_bound"]))
        lines.append("static constexpr int WEC_%s_NTEMPS = %d;" % (cv["name"], nt))
        lines.append("static constexpr int WEC_%s_G = %d;" % (cv["name"], cv["G"]))
        lines.append("")
    return "\n".join(lines) + "\n", summary


def main():
    text, summary = header()
    if "--check" in sys.argv:
        cur = open(OUT).read() if os.path.exists(OUT) else ""
        if cur != text:
            print("wec_programs.cuh is stale: run python tools/gen_wec.py", file=sys.stderr)
            sys.exit(1)
        return
    with open(OUT, "w") as f:
        f.write(text)
    for row in summary:
        print("%-8s %-6s rows %3d  product %2d  linear %2d  reduction %2d  temps %3d  max bound %5d p" % row)


if __name__ == "__main__":
    main()
