"""CPU checks of the C ABI (no compute calls): the library builds/loads, exports every symbol that
include/change3d_b200.h declares, the ctypes structures match the C layout, and bad arguments are
rejected with a status code before anything touches a device."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "change3d_b200.h")


@pytest.fixture(scope="module")
def lib():
    from change3d_b200 import _lib, build
    if not os.path.isfile(_lib.LIB_PATH):
        build.build()
    return _lib.load()


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\bint\s+(c3d_\w+)\s*\(", src)))


def test_header_symbols_exported_and_bound(lib):
    from change3d_b200 import _lib
    names = _declared()
    assert "c3d_pw_gemm" in names and "c3d_adam_step" in names and len(names) >= 17
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported by the library"
    bound = set(_lib.SIGNATURES) | {"c3d_version"}
    assert set(names) == bound, (set(names) ^ bound)
    assert lib.c3d_version() == 1


def test_ctypes_structs_match_c_layout():
    from change3d_b200 import _lib
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "change3d_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(c3d_operand), sizeof(c3d_gemm_desc), sizeof(c3d_wgrad_desc),
         offsetof(c3d_operand, img_stride), offsetof(c3d_operand, seg0), offsetof(c3d_gemm_desc, W),
         offsetof(c3d_gemm_desc, rows_per_sample), offsetof(c3d_wgrad_desc, dW), offsetof(c3d_gemm_desc, flags),
         sizeof(c3d_attn_desc), offsetof(c3d_attn_desc, keep_scale), offsetof(c3d_attn_desc, causal));
  return 0;
}'''
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "t.c"), os.path.join(d, "t")
        open(src, "w").write(prog)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        got = [int(v) for v in subprocess.check_output([exe]).split()]
    want = [C.sizeof(_lib.Operand), C.sizeof(_lib.GemmDesc), C.sizeof(_lib.WgradDesc),
            _lib.Operand.img_stride.offset, _lib.Operand.seg0.offset, _lib.GemmDesc.W.offset,
            _lib.GemmDesc.rows_per_sample.offset, _lib.WgradDesc.dW.offset, _lib.GemmDesc.flags.offset,
            C.sizeof(_lib.AttnDesc), _lib.AttnDesc.keep_scale.offset, _lib.AttnDesc.causal.offset]
    assert got == want


def test_bad_arguments_return_status_without_device(lib):
    from change3d_b200 import _lib
    assert lib.c3d_pw_gemm(None, None) == 1
    assert lib.c3d_pw_wgrad(None, None) == 1
    d = _lib.GemmDesc()                      # all-zero descriptor: rejected by validation
    assert lib.c3d_pw_gemm(C.byref(d), None) == 1
    assert lib.c3d_bn_finalize(None, 0, 0, None, None, None, None, 0, 0, 0.1, 1e-5, 1, None, None) == 1
    assert lib.c3d_dw_conv_fwd(None, None, None, None, None, 1, 3, 8, 8, 24, 24, 1, None) == 1
    assert lib.c3d_dw_conv_fwd(1, 1, 1, 1, None, 1, 7, 8, 8, 24, 24, 1, None) == 1      # T out of range
    assert lib.c3d_adam_step(None, None, None, None, 0, 1e-3, 0.9, 0.99, 1e-8, 0.0, 1, 1.0, 0.0, None) == 1
    assert lib.c3d_stem_fwd(None, None, None, None, None, None, None, 1, 3, 8, 8, None) == 1
    assert lib.c3d_stem_bwd(None, None, None, None, None, None, None, None, None, None, None, None, 1, 3, 8, 8, 1, None) == 1
    # ReLU mask recomputed from y_c (out = NULL) is only defined without a normalised shortcut
    assert lib.c3d_relu_bwd_stats(1, None, 1, 1, 1, 1, None, 1, 1, 64, 24, None) == 1
    assert lib.c3d_relu_bwd_stats(1, None, 1, 1, None, None, None, None, None, 64, 24, None) == 1     # no statistics buffer
    assert lib.c3d_attention_fwd(None, None) == 1
    ad = _lib.AttnDesc()
    ad.q = ad.k = ad.v = ad.o = 1
    ad.B, ad.nh, ad.hd, ad.Lq, ad.Lk = 1, 8, 24, 65, 256          # Lq over the kernel's tile limit
    assert lib.c3d_attention_fwd(C.byref(ad), None) == 1
    assert lib.c3d_augment_pairs(None, 0, None, None, 1, 8, 8, 8, 8, 1, 0, 0.5, 0.5, None, None, None, None) == 1


def test_product_path_fails_loudly_without_cuda():
    import argparse
    import contextlib
    import io
    import torch
    from change3d_b200.model.trainer import Trainer
    from change3d_b200.model.x3d import create_x3d
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    a = argparse.Namespace(num_perception_frame=1, num_class=1, in_height=32, in_width=32, dataset="LEVIR-CD",
                           pretrained="/nonexistent")
    with contextlib.redirect_stdout(io.StringIO()):
        t = Trainer(a)
    with pytest.raises(RuntimeError, match="no CPU"):
        t.update_bcd(torch.zeros(1, 3, 32, 32), torch.zeros(1, 3, 32, 32))
    net = create_x3d(input_clip_length=3, depth_factor=5.0)
    with pytest.raises(RuntimeError, match="no CPU"):
        net.blocks[1](torch.zeros(1, 24, 3, 8, 8))


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    """No silent fallback: with the shared library absent, loading it — and therefore every op wrapper — raises."""
    from change3d_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "libchange3d_b200.so"))
    with pytest.raises(RuntimeError, match="no CPU/eager fallback"):
        _lib.load()
    import torch
    from change3d_b200 import losses
    if not torch.cuda.is_available():
        return
    with pytest.raises(RuntimeError, match="no CPU/eager fallback"):      # on a GPU box: the op itself must raise too
        losses.bce_dice_loss(torch.rand(4, device="cuda"), torch.zeros(4, device="cuda"))
