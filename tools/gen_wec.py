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

Instruction word: last << 31 | op << 24 | dst << 16 | a << 8 | b.  A program is a sequence of ROWS of G words (one per
lane); a lane executes its words in order and the group synchronises after every row whose words carry the `last`
bit, so that a lane can run a chain of dependent linear instructions on its own results without a barrier in
between.  Slot byte: 0..15 = X (the point updated in place), 16..31 = Y (the point added), 32.. = T (temporaries; the
first ones carry values between the two halves of an addition).  One slot = one base-field element (ten u32 words,
Montgomery form).

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

OP_NOP, OP_MUL, OP_ADD, OP_SUB, OP_DBL, OP_NEG, OP_CPY, OP_MULK, OP_INV, OP_SUBD = range(10)
OP_NAMES = ["nop", "mul", "add", "sub", "dbl", "neg", "cpy", "mulk", "inv", "subd"]
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
        c0 = v0 + ((a[1] + a[2]) * (b[1] + b[2]) - v1 - v2).mulk(nr)
        c1 = (a[0] + a[1]) * (b[0] + b[1]) - v0 - v1 + v2.mulk(nr)
        c2 = (a[0] + a[2]) * (b[0] + b[2]) - v0 - v2 + v1
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
        return E([s0 + s3.mulk(nr), s1 + s4.mulk(nr), s1 + s2 + s3 - s0 - s4], nr)

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


# ---- scheduling ------------------------------------------------------------------------------------------------
def schedule(g, outs, G):
    """List scheduling by dependency level: a step holds up to G mutually independent instructions, all linear or all
    products / inversions (a product step costs a full Montgomery product whatever the number of lanes that multiply,
    so products wait until no linear instruction is ready and then go together, deepest remaining chain first).
    Returns a list of steps (kind, lanes), lanes = list of one-instruction chains."""
    nodes = g.nodes
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
    users = {i: [] for i in needed}
    for i in needed:
        n = nodes[i]
        for o in (n.a, n.b):
            if o is not None:
                users[o.id].append(i)
    height = {}

    def h(i):
        if i not in height:
            n = nodes[i]
            w = 10 if n.op in ("mul", "inv") else 1
            height[i] = w + max([h(u) for u in users[i]], default=0)
        return height[i]

    is_prod = lambda i: nodes[i].op in ("mul", "inv")
    done = {i for i in needed if nodes[i].op == "in"}
    todo = [i for i in sorted(needed) if nodes[i].op != "in"]
    steps = []
    ready = lambda i: all(o is None or o.id in done for o in (nodes[i].a, nodes[i].b))
    while todo:
        lin = sorted((i for i in todo if not is_prod(i) and ready(i)), key=lambda i: -h(i))
        if lin:
            pick, kind = lin[:G], "lin"
        else:
            pick, kind = sorted((i for i in todo if is_prod(i) and ready(i)), key=lambda i: -h(i))[:G], "mul"
        assert pick, "scheduler stuck"
        steps.append((kind, [[i] for i in pick]))
        done.update(pick)
        todo = [i for i in todo if i not in done]
    return steps


def allocate(g, outs, steps, G):
    """Slots: inputs are pinned; an output goes straight to its slot when the value living there is dead, else to a
    temporary with a copy at the end; a temporary is reused after the STEP of its last use (a slot is never read and
    written by different lanes inside one step).  Returns rows: list of (row of G instruction tuples | None, last)."""
    nodes = g.nodes
    last_use = {}
    for s, (_, lanes) in enumerate(steps):
        for lane in lanes:
            for i in lane:
                n = nodes[i]
                for o in (n.a, n.b):
                    if o is not None:
                        last_use[o.id] = s
    out_of = {}
    for n, slot in outs:
        out_of.setdefault(n.id, []).append(slot)
    out_slots = {slot for _, slot in outs}
    pinned_until = {}
    for n in nodes:
        if n.op == "in":
            pinned_until[n.pin] = max(pinned_until.get(n.pin, -1), last_use.get(n.id, -1))
    loc = {n.id: n.pin for n in nodes if n.op == "in"}
    t_pinned = {n.pin[1] for n in nodes if n.op == "in" and n.pin[0] == REG_T}
    t_pinned |= {s[1] for s in out_slots if s[0] == REG_T}
    free_t = [i for i in range(MAX_T) if i not in t_pinned]
    release = {}
    final_copies = []
    max_t = max(t_pinned, default=-1)
    rows = []
    for s, (kind, lanes) in enumerate(steps):
        depth = max(len(lane) for lane in lanes)
        step_rows = [[None] * G for _ in range(depth)]
        for li, lane in enumerate(lanes):
            for ri, i in enumerate(lane):
                n = nodes[i]
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
                step_rows[ri][li] = (n.op, dst, loc[n.a.id] if n.a is not None else (0, 0),
                                     loc[n.b.id] if n.b is not None else (0, 0), n.k)
        for ri, r in enumerate(step_rows):
            rows.append((r, ri == depth - 1))
        for t in release.pop(s, []):
            free_t.append(t)
        free_t.sort()
    for n, slot in outs:
        if loc[n.id] != slot and (slot, n.id) not in final_copies:
            final_copies.append((slot, n.id))
    if final_copies:
        srcs = {loc[i] for _, i in final_copies}
        dsts = [s for s, _ in final_copies]
        assert not (srcs & set(dsts)), "final copies alias"
        for c in range(0, len(final_copies), G):
            chunk = [("cpy", slot, loc[i], (0, 0), 0) for slot, i in final_copies[c:c + G]]
            rows.append((chunk + [None] * (G - len(chunk)), True))
    return rows, max_t + 1


def enc_slot(s):
    base = {REG_X: X_BASE, REG_Y: Y_BASE, REG_T: T_BASE}[s[0]]
    lim = {REG_X: 16, REG_Y: 16, REG_T: MAX_T}[s[0]]
    assert 0 <= s[1] < lim
    return base + s[1]


def encode(rows, G):
    words = []
    opc = {n: i for i, n in enumerate(OP_NAMES)}
    for row, last in rows:
        assert len(row) == G
        for ins in row:
            w = 0
            if ins is not None:
                op, d, a, b, k = ins
                bb = k if op == "mulk" else enc_slot(b)
                assert 0 <= bb < 256
                w = (opc[op] << 24) | (enc_slot(d) << 16) | (enc_slot(a) << 8) | bb
            words.append(w | (0x80000000 if last else 0))
    return words


def build(curve, prog, G=None):
    G = G or CURVES[curve]["G"]
    g, outs = trace(curve, prog)
    steps = schedule(g, outs, G)
    rows, nt = allocate(g, outs, steps, G)
    return dict(curve=curve, prog=prog, G=G, rows=rows, words=encode(rows, G), nrows=len(rows), ntemps=nt,
                nsteps=sum(1 for _, last in rows if last),
                mul_steps=sum(1 for kind, _ in steps if kind == "mul"),
                lin_rows=sum(max(len(l) for l in lanes) for kind, lanes in steps if kind == "lin"))


# ---- Python interpreter (tests) ----------------------------------------------------------------------------------
def interpret(built, p, X, Y, T=None):
    """Run a schedule on Python integers mod p.  X, Y, T: lists of ints (modified in place).  Also checks the
    hazards the GPU interpreter relies on: inside one step (rows up to a `last` row) no slot written by a lane is read
    or written by another lane."""
    G = built["G"]
    T = T if T is not None else [0] * MAX_T
    X += [0] * (16 - len(X))
    Y += [0] * (16 - len(Y))
    mem = X + Y + T

    words = built["words"]
    step_reads = [set() for _ in range(G)]
    step_writes = [set() for _ in range(G)]
    for r in range(built["nrows"]):
        last = False
        for lane in range(G):
            w = words[r * G + lane]
            last = bool(w >> 31)
            op, d, a, b = (w >> 24) & 127, (w >> 16) & 255, (w >> 8) & 255, w & 255
            if op == OP_NOP:
                continue
            av = mem[a]
            step_reads[lane].add(a)
            if op in (OP_MUL, OP_ADD, OP_SUB, OP_SUBD):
                bv = mem[b]
                step_reads[lane].add(b)
            if op == OP_MUL:
                res = av * bv % p
            elif op == OP_ADD:
                res = (av + bv) % p
            elif op == OP_SUB:
                res = (av - bv) % p
            elif op == OP_DBL:
                res = 2 * av % p
            elif op == OP_NEG:
                res = (-av) % p
            elif op == OP_CPY:
                res = av
            elif op == OP_MULK:
                res = av * b % p
            elif op == OP_INV:
                res = pow(av, p - 2, p)
            elif op == OP_SUBD:
                res = (av - 2 * bv) % p
            else:
                raise ValueError(op)
            # lanes run at their own pace inside a step: executing them one after the other is one legal order, and
            # the hazard check below makes every order equivalent
            mem[d] = res
            step_writes[lane].add(d)
        if last:
            for l1 in range(G):
                for l2 in range(G):
                    if l1 != l2:
                        assert not (step_writes[l1] & (step_reads[l2] | step_writes[l2])), "cross-lane hazard in a step"
            step_reads = [set() for _ in range(G)]
            step_writes = [set() for _ in range(G)]
    X[:] = mem[0:16]
    Y[:] = mem[16:32]
    T[:] = mem[32:]
    return X, Y, T


# ---- header ------------------------------------------------------------------------------------------------------
def header():
    lines = ["// GENERATED by tools/gen_wec.py -- do not edit (python tools/gen_wec.py --check verifies it is current).",
             "// Lane schedules of the warp-cooperative group law (wec.cuh): one u32 per lane and row,",
             "// last << 31 | op << 24 | dst << 16 | a << 8 | b; slots 0..15 X, 16..31 Y, 32.. T; ops: " +
             ", ".join("%d %s" % (i, n) for i, n in enumerate(OP_NAMES)) + ".",
             "#pragma once", "#include \"prims.cuh\"", ""]
    summary = []
    for curve in sorted(CURVES):
        cv = CURVES[curve]
        nt = 0
        for prog in PROGRAMS:
            b = build(curve, prog)
            nt = max(nt, b["ntemps"])
            name = "WEC_%s_%s" % (cv["name"], prog.upper())
            lines.append("// %s %s: %d rows in %d steps (%d product steps, %d linear rows), %d temporaries, %d lanes" %
                         (cv["name"], prog, b["nrows"], b["nsteps"], b["mul_steps"], b["lin_rows"], b["ntemps"], b["G"]))
            lines.append("__device__ const u32 %s[%d] = {" % (name, len(b["words"])))
            ws = b["words"]
            for c in range(0, len(ws), 8):
                lines.append("  " + ", ".join("0x%08xu" % w for w in ws[c:c + 8]) + ",")
            lines.append("};")
            lines.append("static constexpr int %s_ROWS = %d;" % (name, b["nrows"]))
            summary.append((cv["name"], prog, b["nrows"], b["nsteps"], b["mul_steps"], b["lin_rows"], b["ntemps"]))
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
        print("%-8s %-6s rows %3d  steps %2d  product steps %2d  linear rows %2d  temps %2d" % row)


if __name__ == "__main__":
    main()
