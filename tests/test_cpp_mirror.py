"""The C++ host mirror of the reference's SNARK interface (include/pcdgpu_snark.hpp), driven by a small
C++ program the way the reference's tests drive Groth16::prove.  CPU: it compiles, links against
libpcdgpu.so and refuses to run without a GPU (exit code 2: no CPU fallback).  GPU: byte parity with
the golden vectors."""
import os
import subprocess

import numpy as np
import pytest

import codec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "cpp", "_build")
EXE = os.path.join(BUILD, "snark_mirror_test")


def _build():
    lib = os.path.join(ROOT, "pcd_b200", "libpcdgpu.so")
    if not os.path.exists(lib):
        pytest.skip("libpcdgpu.so not built yet")
    src = os.path.join(ROOT, "tests", "cpp", "snark_mirror_test.cpp")
    hdr = os.path.join(ROOT, "include", "pcdgpu_snark.hpp")
    if not os.path.exists(EXE) or max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(EXE):
        os.makedirs(BUILD, exist_ok=True)
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-pthread", "-I", os.path.join(ROOT, "include"), src, "-o", EXE,
                               "-L", os.path.join(ROOT, "pcd_b200"), "-lpcdgpu",
                               "-Wl,-rpath," + os.path.join(ROOT, "pcd_b200"),
                               "-Wl,-rpath-link,/usr/local/cuda/lib64", "-L/usr/local/cuda/lib64", "-lcudart"])
    return EXE


def _fixture(case, path):
    pid = case["pairing"]
    g1, g2 = codec.G1_OF[pid], codec.G2_OF[pid]
    words = [pid, case["m"], case["num_inputs"], case["num_witness"]]
    for k in "ABC":
        ptr, col, val = codec.csr_from_golden(case[k])
        for i in range(case["m"]):
            lo, hi = int(ptr[i]), int(ptr[i + 1])
            words.append(hi - lo)
            for e in range(lo, hi):
                words += [int(x) for x in val[e]] + [int(col[e])]
    words += [int(x) for x in codec.hex_to_u64(case["z"])]
    words += [int(x) for x in codec.hex_to_u64(case["r"])] + [int(x) for x in codec.hex_to_u64(case["s"])]
    pk = case["pk"]
    for k in ("alpha_g1", "beta_g1", "delta_g1", "beta_g2", "delta_g2", "a_query", "b_g1_query", "b_g2_query"):
        words += [int(x) for x in codec.hex_to_u64(pk[k])]
    hq = codec.hex_to_u64(pk["h_query"], codec.POINT_LIMBS[g1])
    words.append(hq.shape[0])
    words += [int(x) for x in hq.reshape(-1)]
    words += [int(x) for x in codec.hex_to_u64(pk["l_query"])]
    np.array(words, dtype=np.uint64).tofile(path)
    del g2


def _gm17_fixture(case, path):
    pid = case["pairing"]
    g1 = codec.G1_OF[pid]
    words = [pid | (1 << 8), case["m"], case["num_inputs"], case["num_witness"]]
    for k in "ABC":
        ptr, col, val = codec.csr_from_golden(case[k])
        for i in range(case["m"]):
            lo, hi = int(ptr[i]), int(ptr[i + 1])
            words.append(hi - lo)
            for e in range(lo, hi):
                words += [int(x) for x in val[e]] + [int(col[e])]
    words += [int(x) for x in codec.hex_to_u64(case["z"])]
    for k in ("d1", "d2", "r"):
        words += [int(x) for x in codec.hex_to_u64(case[k])]
    pk = case["pk"]
    nsap = codec.hex_to_u64(pk["a_query"], codec.POINT_LIMBS[g1]).shape[0]
    words += [nsap, case["domain_size"] + 1]
    for k in ("a_query", "b_query", "c_query_1", "c_query_2", "g_gamma2_z_t", "g_gamma_z", "h_gamma_z", "g_ab_gamma_z",
              "g_gamma2_z2"):
        words += [int(x) for x in codec.hex_to_u64(pk[k])]
    np.array(words, dtype=np.uint64).tofile(path)


def test_cpp_mirror_builds_and_refuses_cpu(tmp_path):
    import torch
    exe = _build()
    case = codec.load("groth16")[0]
    fx, out = str(tmp_path / "fx.bin"), str(tmp_path / "out.bin")
    _fixture(case, fx)
    rc = subprocess.call([exe, fx, out])
    if torch.cuda.is_available():
        assert rc == 0
    else:
        assert rc == 2  # PCDGPU_E_NODEVICE surfaced through Groth16::index: no CPU fallback
    _gm17_fixture(codec.load("gm17")[0], fx)
    rc = subprocess.call([exe, fx, out])
    assert rc == (0 if torch.cuda.is_available() else 2)


@pytest.mark.gpu
def test_cpp_mirror_matches_golden(tmp_path):
    exe = _build()
    for i, case in enumerate(codec.load("groth16")):
        fx, out = str(tmp_path / ("fx%d.bin" % i)), str(tmp_path / ("out%d.bin" % i))
        _fixture(case, fx)
        assert subprocess.call([exe, fx, out]) == 0
        blob = open(out, "rb").read()
        n_aff = len(case["proof_affine"]) // 2
        assert blob[:n_aff].hex() == case["proof_affine"]
        assert blob[n_aff:].hex() == case["proof_bytes"]


@pytest.mark.gpu
def test_cpp_mirror_gm17_matches_golden(tmp_path):
    exe = _build()
    for i, case in enumerate(codec.load("gm17")):
        fx, out = str(tmp_path / ("gfx%d.bin" % i)), str(tmp_path / ("gout%d.bin" % i))
        _gm17_fixture(case, fx)
        assert subprocess.call([exe, fx, out]) == 0
        blob = open(out, "rb").read()
        n_aff = len(case["proof_affine"]) // 2
        assert blob[:n_aff].hex() == case["proof_affine"]
        assert blob[n_aff:].hex() == case["proof_bytes"]
