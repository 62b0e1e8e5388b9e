"""The C-ABI shared library builds for sm_100a, loads on a CPU-only box and exports every symbol that
include/samcon_b200.h declares (no compute calls here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from samcon_b200 import build
    if not os.path.exists(os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")) and not os.path.exists(build.LIB):
        pytest.skip("no nvcc and no prebuilt library")
    return build.build()


def test_header_symbols_exported(lib_path):
    hdr = open(os.path.join(ROOT, "include", "samcon_b200.h")).read()
    declared = set(re.findall(r"^(?:const char \*|int64_t |int )\s*(sc_\w+)\(", hdr, flags=re.M))
    from samcon_b200 import abi
    assert declared == set(abi.SYMBOLS) and len(declared) >= 11
    lib = C.CDLL(lib_path)
    for s in declared:
        assert hasattr(lib, s), s


def test_ctypes_struct_matches_header_layout(character):
    """ScModel mirrors `struct sc_model` field for field (names and order)."""
    from samcon_b200 import abi
    hdr = open(os.path.join(ROOT, "include", "samcon_b200.h")).read()
    body = hdr[hdr.index("typedef struct sc_model {"):hdr.index("} sc_model;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";")[:-1]:
        decl = decl.split("{")[-1]
        for part in decl.split(","):
            m = re.search(r"(\w+)\s*(\[\d+\])?\s*$", part.strip())
            if m:
                names.append(m.group(1))
    assert names == [f[0] for f in abi.ScModel._fields_]
    m, keep = abi.make_model(character)
    assert m.nb == 18 and m.nlev == 7 and m.ncand == 73 and m.ncp == 20 and m.nee == 4
    assert abs(m.dt - 1e-3) < 1e-9 and m.num_substeps == 2 and m.num_solver_iters == 10


def test_pool_refuses_without_gpu(character):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from samcon_b200.pool import GpuSamplePool
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        GpuSamplePool(character, 0)


def test_sc_create_validates_the_model_before_touching_cuda(lib_path, character):
    """Error behaviour of the C ABI (no compute, works on a box without a GPU): a malformed model is refused with
    SC_EINVAL and a message; every entry point refuses a null handle."""
    from samcon_b200 import abi
    lib = abi.load()
    m, keep = abi.make_model(character)
    h = C.c_void_p()
    m.max_contacts = 99
    rc = lib.sc_create(C.byref(h), C.byref(m), 0)
    assert rc == -1 and b"max_contacts" in lib.sc_last_error(h)
    assert lib.sc_destroy(h) == 0
    m, keep = abi.make_model(character)
    m.nlev = 1                                                    # level table no longer covers the 18 bodies
    h = C.c_void_p()
    assert lib.sc_create(C.byref(h), C.byref(m), 0) == -1 and len(lib.sc_last_error(h)) > 0
    lib.sc_destroy(h)
    m, keep = abi.make_model(character)
    m.npair = 4000                                                # more primitive pairs than SC_MAX_PAIRS
    h = C.c_void_p()
    assert lib.sc_create(C.byref(h), C.byref(m), 0) == -1 and b"npair" in lib.sc_last_error(h)
    lib.sc_destroy(h)
    assert lib.sc_create(None, C.byref(m), 0) == -1
    assert lib.sc_destroy(None) == -1 and lib.sc_launch_count(None) == 0
    assert lib.sc_window(None, None, None, 1, None, 1, None, None, 1, None, 1, 1, None, None, None, None, None, None) == -1
    assert lib.sc_select(None, None, None, 1, 1, None, None, 0, None) == -1
    assert lib.sc_last_error(None) == b"null handle"
