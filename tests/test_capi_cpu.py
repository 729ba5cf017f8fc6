"""The C-ABI library loads and exports every symbol include/rtf_b200.h declares (no GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    import recommend_tf2_b200 as pkg
    return pkg


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "rtf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\bint\s+(rtf_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(built):
    names = _declared_symbols()
    assert "rtf_embed_fwd" in names and "rtf_embed_bwd" in names
    handle = ctypes.CDLL(built._lib.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), f"{n} declared in rtf_b200.h but not exported"


def test_python_binding_covers_header(built):
    assert sorted(built._lib.SIGNATURES) == _declared_symbols()


def test_version_and_arg_errors_without_gpu(built):
    lib = built.lib()
    arch = ctypes.c_int(0)
    assert lib.rtf_version(ctypes.byref(arch)) >= 1 and arch.value == 100
    # argument validation happens before any CUDA call
    assert lib.rtf_embed_fwd(None, None, None, 0, None, 0, 0, 1, 0, 0, 0, 0, None, 0, None, None) == -1
    nbytes = ctypes.c_size_t(0)
    assert lib.rtf_embed_bwd_workspace(1 << 20, 128, ctypes.byref(nbytes)) == 0
    assert nbytes.value > (1 << 20) * 16
    # dense GEMM: NULL operands -> RTF_E_ARG, a leading dimension that breaks the 16-byte rows -> RTF_E_ALIGN
    for fn in (lib.rtf_dense_gemm_nn, lib.rtf_dense_gemm_nt, lib.rtf_dense_gemm_tn):
        assert fn(None, 8, 0, None, 8, 0, None, 0, None, 8, 0, 4, 8, 8, 1, None, 0, None) == -1
        assert fn(0x1000, 6, 0, 0x2000, 8, 0, None, 0, 0x3000, 8, 0, 4, 8, 8, 1, None, 0, None) == -2
    # peer-memory interaction: missing pointer tables -> RTF_E_ARG
    assert lib.rtf_embed_dot_peer_fwd(None, None, 0, 2, 0, None, 3, 8, None, 0, 4, 3, 1, None, 8, None, 16,
                                      16, None, 0, None, None) == -1
    assert lib.rtf_embed_dot_peer_bwd(None, 0, None, 3, 8, None, 0, 4, 3, 1, None, 8, None, 16, None, 8,
                                      None, None, 2, 0, 0, None) == -1


def test_product_path_has_no_cpu_fallback(built):
    import torch
    with pytest.raises(built.RtfError):
        built.embed_fwd([torch.zeros(4, 4)], torch.zeros((2, 1), dtype=torch.int32))


def test_product_package_never_imports_oracle():
    pkg_dir = os.path.join(ROOT, "recommend-tf2.0_b200")
    for dp, _, fns in os.walk(pkg_dir):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, fn)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, fn


def test_tf_shim_source_type_checks_against_the_c_header(tmp_path):
    """tf_ops/rtf_tf_ops.cc (the tf.load_op_library shim) cannot be built here (no
    TensorFlow headers).  It is TYPE-CHECKED instead against a declaration-only stand-in for the
    TensorFlow names it uses (tests/tf_stub, test infrastructure) and the REAL include/rtf_b200.h:
    every call into librtf_b200 must match the C-ABI in argument count and types; a deliberately
    broken call must be rejected (the check has teeth)."""
    import shutil
    import subprocess
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = os.path.join(root, "tf_ops", "rtf_tf_ops.cc")
    cmd = [gxx, "-std=c++17", "-fsyntax-only", "-I" + os.path.join(root, "tests", "tf_stub"),
           "-I" + os.path.join(root, "include")]
    ok = subprocess.run(cmd + [src], capture_output=True, text=True)
    assert ok.returncode == 0, ok.stderr[-2000:]
    text = open(src).read()
    for name in ("RtfEmbedFwd", "RtfEmbedBwdAdam", "RtfEmbedDotFwd", "RtfEmbedDotBwd", "RtfBce"):
        assert f'REGISTER_OP("{name}")' in text and f'Name("{name}")' in text
    broken = tmp_path / "broken.cc"
    needle = "ws.flat<uint8>().data(), StreamOf(ctx));"
    assert needle in text
    broken.write_text(text.replace(needle, "StreamOf(ctx));"))      # one argument short
    bad = subprocess.run(cmd + [str(broken)], capture_output=True, text=True)
    assert bad.returncode != 0 and "rtf_bce_fwd" in bad.stderr
