"""ctypes binding of oracle/c/liboracle.so (the C++ CPU oracle / CPU baseline arm).

TEST INFRASTRUCTURE ONLY -- see oracle/c/pcd_oracle.cpp.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs import this module.  All buffers are numpy arrays
in the ABI encodings of include/pcdgpu.h (uint64 little-endian limbs).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "c", "liboracle.so")

# curve ids (same as include/pcdgpu.h)
MNT4_G1, MNT4_G2, MNT6_G1, MNT6_G2 = 0, 1, 2, 3
#: u64 limbs per affine point (x || y)
POINT_LIMBS = {0: 10, 1: 20, 2: 10, 3: 30}
#: scalar field id (0 = r4, 1 = q4) of each curve, and coordinate base field
SCALAR_FIELD = {0: 0, 1: 0, 2: 1, 3: 1}


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "c", "pcd_oracle.cpp")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", os.path.join(HERE, "c"), "-B", "liboracle.so"],
                              stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.orc_hw_threads.restype = ctypes.c_int
        _lib.orc_witness_map.restype = ctypes.c_int
        _lib.orc_groth16_prove.restype = ctypes.c_int
        _lib.orc_serialize_proof.restype = ctypes.c_int
        _lib.orc_on_curve.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def hw_threads() -> int:
    return lib().orc_hw_threads()


def field_op(field: int, op: int, a: np.ndarray, b: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    b = np.ascontiguousarray(b, dtype=np.uint64)
    out = np.zeros_like(a)
    lib().orc_field_op(field, op, _p(a), _p(b), _p(out))
    return out


def ntt(field: int, data: np.ndarray, inverse: bool, coset: bool, threads: int = 1) -> np.ndarray:
    d = np.array(data, dtype=np.uint64, copy=True).reshape(-1, 5)
    n = d.shape[0]
    log_n = n.bit_length() - 1
    assert 1 << log_n == n
    lib().orc_ntt(field, _p(d), log_n, int(inverse), int(coset), threads)
    return d


def dft_at(field: int, data: np.ndarray, indices, coset: bool = False, threads: int = 0) -> np.ndarray:
    """out[t] = sum_j data[j] (g^coset omega_n^indices[t])^j: the transform's definition at a few output indices"""
    d = np.ascontiguousarray(data, dtype=np.uint64).reshape(-1, 5)
    n = d.shape[0]
    log_n = n.bit_length() - 1
    assert 1 << log_n == n
    idx = np.ascontiguousarray(indices, dtype=np.uint64)
    out = np.zeros((len(idx), 5), dtype=np.uint64)
    lib().orc_dft_at(field, _p(d), log_n, int(coset), _p(idx), ctypes.c_size_t(len(idx)), _p(out), threads or hw_threads())
    return out


def domain_size(field: int, min_size: int):
    """GeneralEvaluationDomain::new(min_size) -> (n, a, b) with n = 7^a 2^b, or None."""
    a, b = ctypes.c_int(), ctypes.c_int()
    lib().orc_domain_size.restype = ctypes.c_size_t
    n = lib().orc_domain_size(field, ctypes.c_size_t(min_size), ctypes.byref(a), ctypes.byref(b))
    return (n, a.value, b.value) if n else None


def ntt_general(field: int, data: np.ndarray, a: int, b: int, inverse: bool, coset: bool, threads: int = 1) -> np.ndarray:
    d = np.array(data, dtype=np.uint64, copy=True).reshape(-1, 5)
    assert d.shape[0] == (7 ** a) << b
    lib().orc_ntt_general(field, _p(d), a, b, int(inverse), int(coset), threads)
    return d


def msm(curve: int, bases: np.ndarray, scalars: np.ndarray, threads: int = 1, c: int = 0) -> np.ndarray:
    bases = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, POINT_LIMBS[curve])
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 5)
    n = min(bases.shape[0], scalars.shape[0])
    out = np.zeros(POINT_LIMBS[curve], dtype=np.uint64)
    lib().orc_msm(curve, _p(bases), _p(scalars), ctypes.c_size_t(n), _p(out), threads, c)
    return out


def fixed_base_mul(curve: int, base: np.ndarray, scalars: np.ndarray, threads: int = 0) -> np.ndarray:
    base = np.ascontiguousarray(base, dtype=np.uint64)
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 5)
    n = scalars.shape[0]
    out = np.zeros((n, POINT_LIMBS[curve]), dtype=np.uint64)
    lib().orc_fixed_base_mul(curve, _p(base), _p(scalars), ctypes.c_size_t(n), _p(out), threads or hw_threads())
    return out


def point_sum(curve: int, pts: np.ndarray) -> np.ndarray:
    pts = np.ascontiguousarray(pts, dtype=np.uint64).reshape(-1, POINT_LIMBS[curve])
    out = np.zeros(POINT_LIMBS[curve], dtype=np.uint64)
    lib().orc_point_sum(curve, _p(pts), ctypes.c_size_t(pts.shape[0]), _p(out))
    return out


def on_curve(curve: int, pts: np.ndarray, b_coeff: np.ndarray) -> bool:
    pts = np.ascontiguousarray(pts, dtype=np.uint64).reshape(-1, POINT_LIMBS[curve])
    b_coeff = np.ascontiguousarray(b_coeff, dtype=np.uint64)
    return bool(lib().orc_on_curve(curve, _p(pts), ctypes.c_size_t(pts.shape[0]), _p(b_coeff)))


def from_mont(field: int, a: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 5)
    out = np.zeros_like(a)
    lib().orc_from_mont(field, _p(a), _p(out), ctypes.c_size_t(a.shape[0]))
    return out


def to_mont(field: int, a: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 5)
    out = np.zeros_like(a)
    lib().orc_to_mont(field, _p(a), _p(out), ctypes.c_size_t(a.shape[0]))
    return out


def _csr_args(csr):
    ptr, col, val = csr
    ptr = np.ascontiguousarray(ptr, dtype=np.uint32)
    col = np.ascontiguousarray(col, dtype=np.uint32)
    val = np.ascontiguousarray(val, dtype=np.uint64)
    return (ptr, col, val), [_p(ptr), _p(col), _p(val)]


def witness_map(pairing: int, A, B, C, m: int, num_inputs: int, z: np.ndarray, threads: int = 1) -> np.ndarray:
    keep, args = [], []
    for M in (A, B, C):
        k, a = _csr_args(M)
        keep.append(k)
        args += a
    z = np.ascontiguousarray(z, dtype=np.uint64).reshape(-1, 5)
    dom = domain_size(pairing, m + num_inputs)
    if dom is None:
        raise ValueError("domain too large for the field")
    n = dom[0]
    h = np.zeros((n, 5), dtype=np.uint64)
    rc = lib().orc_witness_map(pairing, *args, ctypes.c_size_t(m), ctypes.c_size_t(num_inputs), _p(z), _p(h), threads)
    if rc < 0:
        raise ValueError("domain too large for the field")
    return h


PK_FIELDS = ("alpha_g1", "beta_g1", "delta_g1", "beta_g2", "delta_g2", "a_query", "b_g1_query", "b_g2_query",
             "h_query", "l_query")


def groth16_prove(pairing: int, pk: dict, A, B, C, m: int, num_inputs: int, num_witness: int, z: np.ndarray,
                  r: np.ndarray, s: np.ndarray, threads: int = 1) -> np.ndarray:
    """pk: dict of numpy uint64 arrays keyed by PK_FIELDS.  Returns the proof as affine bytes
    A || B || C (uint64 limbs: 10 + 20 + 10 for MNT4, 10 + 30 + 10 for MNT6)."""
    arrs = [np.ascontiguousarray(pk[k], dtype=np.uint64) for k in PK_FIELDS]
    ptrs = (ctypes.c_void_p * 10)(*[a.ctypes.data for a in arrs])
    keep, args = [], []
    for M in (A, B, C):
        k, a = _csr_args(M)
        keep.append(k)
        args += a
    z = np.ascontiguousarray(z, dtype=np.uint64).reshape(-1, 5)
    r = np.ascontiguousarray(r, dtype=np.uint64)
    s = np.ascontiguousarray(s, dtype=np.uint64)
    out = np.zeros(40 if pairing == 0 else 50, dtype=np.uint64)
    rc = lib().orc_groth16_prove(pairing, ptrs, *args, ctypes.c_size_t(m), ctypes.c_size_t(num_inputs),
                                 ctypes.c_size_t(num_witness), _p(z), _p(r), _p(s), _p(out), threads)
    if rc < 0:
        raise ValueError("groth16_prove failed (domain too large)")
    return out


def serialize_proof(pairing: int, proof_affine: np.ndarray) -> bytes:
    proof_affine = np.ascontiguousarray(proof_affine, dtype=np.uint64)
    out = np.zeros(190, dtype=np.uint8)
    n = lib().orc_serialize_proof(pairing, _p(proof_affine), _p(out))
    return out[:n].tobytes()


GM17_PK_FIELDS = ("a_query", "b_query", "c_query_1", "c_query_2", "g_gamma2_z_t", "g_gamma_z", "h_gamma_z",
                  "g_ab_gamma_z", "g_gamma2_z2")


def sap_domain_size(pairing: int, m: int, num_inputs: int):
    dom = domain_size(pairing, 2 * m + 2 * (num_inputs - 1) + 1)
    return dom[0] if dom else None


def sap_witness_map(pairing: int, A, B, C, m: int, num_inputs: int, num_witness: int, z: np.ndarray, d1: np.ndarray,
                    d2: np.ndarray, threads: int = 1):
    """R1CStoSAP::witness_map; d1, d2 plain-integer limbs.  Returns (full SAP assignment, h with n + 1 coefficients)."""
    keep, args = [], []
    for M in (A, B, C):
        k, a = _csr_args(M)
        keep.append(k)
        args += a
    z = np.ascontiguousarray(z, dtype=np.uint64).reshape(-1, 5)
    n = sap_domain_size(pairing, m, num_inputs)
    if n is None:
        raise ValueError("domain too large for the field")
    d1m = to_mont(pairing, np.ascontiguousarray(d1, dtype=np.uint64).reshape(1, 5))
    d2m = to_mont(pairing, np.ascontiguousarray(d2, dtype=np.uint64).reshape(1, 5))
    full = np.zeros((num_inputs + num_witness + m + num_inputs - 1, 5), dtype=np.uint64)
    h = np.zeros((n + 1, 5), dtype=np.uint64)
    rc = lib().orc_sap_witness_map(pairing, *args, ctypes.c_size_t(m), ctypes.c_size_t(num_inputs),
                                   ctypes.c_size_t(num_witness), _p(z), _p(d1m), _p(d2m), _p(full), _p(h), threads)
    if rc < 0:
        raise ValueError("domain too large for the field")
    return full, h


def gm17_prove(pairing: int, pk: dict, A, B, C, m: int, num_inputs: int, num_witness: int, z: np.ndarray,
               d1: np.ndarray, d2: np.ndarray, r: np.ndarray, threads: int = 1) -> np.ndarray:
    """pk: dict of numpy uint64 arrays keyed by GM17_PK_FIELDS; d1, d2, r plain-integer limbs.  Returns the proof as
    affine limbs A || B || C."""
    arrs = [np.ascontiguousarray(pk[k], dtype=np.uint64) for k in GM17_PK_FIELDS]
    ptrs = (ctypes.c_void_p * 9)(*[a.ctypes.data for a in arrs])
    keep, args = [], []
    for M in (A, B, C):
        k, a = _csr_args(M)
        keep.append(k)
        args += a
    z = np.ascontiguousarray(z, dtype=np.uint64).reshape(-1, 5)
    d1, d2, r = (np.ascontiguousarray(x, dtype=np.uint64) for x in (d1, d2, r))
    out = np.zeros(40 if pairing == 0 else 50, dtype=np.uint64)
    rc = lib().orc_gm17_prove(pairing, ptrs, *args, ctypes.c_size_t(m), ctypes.c_size_t(num_inputs),
                              ctypes.c_size_t(num_witness), _p(z), _p(d1), _p(d2), _p(r), _p(out), threads)
    if rc < 0:
        raise ValueError("gm17_prove failed (domain too large)")
    return out
