"""The C-ABI library loads and exports every symbol include/p3r.h declares; without a GPU it fails loudly (no fallback)."""
import importlib
import os
import re

import pytest

lib = importlib.import_module("plonky3-recursion_b200.lib")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "p3r.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(p3r_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    l = lib.load()
    names = header_functions()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(l, n)]
    assert not missing, f"declared in include/p3r.h but not exported: {missing}"
    assert sorted(lib.EXPORTS) == names, "lib.EXPORTS and include/p3r.h disagree"
    assert l.p3r_abi_version() == 1
    assert b"sm_100a" in l.p3r_build_info()


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback_without_gpu():
    with pytest.raises(lib.P3RError) as e:
        lib.Context("koala-bear")
    assert "no usable CUDA device" in str(e.value) or e.value.code == 2


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may touch oracle/ (grep of the product package)."""
    pkg = os.path.join(ROOT, "plonky3-recursion_b200")
    offenders = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"oracle_py|liboracle|orc_[a-z_]+\(", text):
                    offenders.append(os.path.join(dirpath, f))
    assert not offenders, offenders


@pytest.mark.parametrize("field", ["koala-bear", "baby-bear"])
def test_host_hasher_matches_the_oracle_and_the_numpy_restatement(field):
    """p3r_host_hasher (the prover's host-side transcript permutation; host code only, so it runs without a GPU) against two
    independent restatements: the oracle's Poseidon2 and poseidon2_params.permute."""
    import sys
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from common import make_oracle
    p2mod = importlib.import_module("plonky3-recursion_b200.poseidon2_params")
    orc = make_oracle(field)
    F = orc.field
    rng = np.random.default_rng(11)
    st = np.concatenate([F.rand(rng, (64, 16)), np.zeros((1, 16), dtype=np.uint32), np.full((1, 16), F.p - 1, dtype=np.uint32)])
    hh = lib.HostHasher(field)
    got = hh.permute(st)
    assert np.array_equal(got, np.asarray(orc.poseidon2_permute(st), dtype=np.uint32))
    assert np.array_equal(got, np.asarray(p2mod.Poseidon2Params(F.field_id).permute(st.astype(np.uint64)), dtype=np.uint32))
    one = hh.permute(st[3])
    assert one.shape == (1, 16) and np.array_equal(one[0], got[3])
    hh.close()


def test_null_handles_are_refused_not_dereferenced():
    """Every context-taking entry point of the boundary answers a NULL context (or NULL out-pointer) with
    P3R_ERR_INVALID_ARG — checked here without a GPU, the calls never reach CUDA."""
    import ctypes as C
    l = lib.load()
    INVALID = 1
    zero = C.c_uint32(0)
    for name, args in [
        ("p3r_ctx_set_stream_priority", (None, 1)),
        ("p3r_ctx_set_uni_stark", (None, 1)),
        ("p3r_ctx_set_conventions", (None, None)),
        ("p3r_ctx_set_leaf_hasher", (None, None)),
        ("p3r_poseidon2_permute_w", (None, None, None, 0)),
        ("p3r_poseidon2_run_chains", (None, None, None, None)),
        ("p3r_grind", (None, None, None, 0, 4, C.byref(zero))),
        ("p3r_traces_write_rows", (None, None, None, 0, 0, 0, None)),
    ]:
        fn = getattr(l, name)
        fn.restype = C.c_int
        assert fn(*args) == INVALID, name
    out = C.c_void_p()
    assert l.p3r_ctx_create(0, None, None, None, C.byref(out)) == INVALID and not out.value
