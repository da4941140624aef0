"""CPU-side checks of the C-ABI boundary: the in-tree library builds/loads, exports every symbol the
header declares, sizes are sane, errors are reported (not swallowed), and there is no CPU fallback."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "genvox_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gvx_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from genvox_b200 import _native
    lib = _native.load()
    declared = _header_symbols()
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/genvox_b200.h but not exported"
    assert sorted(_native.EXPORTS) == declared            # the ctypes table and the header agree
    assert lib.gvx_abi_version() == 1
    assert os.path.samefile(_native.library_path(), os.path.join(ROOT, "genvox_b200", "lib", "libgenvox_b200.so"))


def test_buffer_sizes_follow_the_shapes():
    import genvox_b200
    from genvox_b200 import _native
    lib = _native.load()
    dec = genvox_b200.Decoder(80, 512, 1024, 256, 1000, 0.5, 0.1, 0.1, 1024, 128, 32, 31)
    d = dec._dims()
    packed = lib.gvx_dec_packed_bytes(C.byref(d))
    n_params = sum(p.numel() for p in dec.parameters())
    assert n_params == 18_255_505
    assert 4 * n_params < packed < 3 * 4 * n_params       # packed = weights + transposed LSTM copies
    s1, s2 = lib.gvx_dec_stash_bytes(C.byref(d), 16, 120, 600), lib.gvx_dec_stash_bytes(C.byref(d), 16, 120, 1200)
    assert 1.9 < s2 / s1 < 2.1
    assert lib.gvx_dec_infer_workspace_bytes(C.byref(d), 64, 150, 1000) > 64 * 1000 * 81 * 4
    assert lib.gvx_dec_stash_bytes(C.byref(d), 0, 120, 600) == 0


def test_invalid_dims_are_reported_not_swallowed():
    from genvox_b200 import _native
    lib = _native.load()
    bad = _native.GvxDims(80, 512, 1024, 1024, 256, 128, 32, 30, 0.1, 0.1)      # even conv kernel
    assert lib.gvx_dec_packed_bytes(C.byref(bad)) == 0
    assert b"odd" in lib.gvx_last_error()
    bad = _native.GvxDims(81, 512, 1024, 1024, 256, 128, 32, 31, 0.1, 0.1)      # n_mels not a multiple of 4
    w = _native.GvxWeights()
    assert lib.gvx_dec_pack_weights(C.byref(bad), C.byref(w), None, None) != 0
    with pytest.raises(RuntimeError, match="multiples of 4"):
        _native.check(1, "gvx_dec_pack_weights")


def test_state_dict_matches_the_reference_decoder_keys():
    import genvox_b200
    from genvox_b200 import _native
    dec = genvox_b200.Decoder(80, 512, 1024, 256, 1000, 0.5, 0.1, 0.1, 1024, 128, 32, 31)
    sd = dec.state_dict()
    assert sorted(sd) == sorted(k for _, k in _native.PARAM_FIELDS)
    assert tuple(sd["attention_rnn.weight_ih"].shape) == (4096, 768)
    assert tuple(sd["decoder_rnn.weight_ih"].shape) == (4096, 1536)
    assert tuple(sd["attention_layer.location_layer.location_conv.conv.weight"].shape) == (32, 2, 31)
    assert tuple(sd["gate_layer.linear_layer.weight"].shape) == (1, 1536)


def test_no_cpu_fallback():
    import genvox_b200
    dec = genvox_b200.Decoder(80, 512, 1024, 256, 1000, 0.5, 0.1, 0.1, 1024, 128, 32, 31)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dec(torch.zeros(2, 5, 512), torch.zeros(2, 80, 3), torch.tensor([5, 4]))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dec.inference(torch.zeros(1, 5, 512))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "genvox_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f


def test_build_stamp_does_not_depend_on_the_checkout_path(tmp_path, monkeypatch):
    """The repo is snapshotted to another directory on the GPU box: the prebuilt library must still count as fresh there
    (a path-dependent stamp made every rank rebuild it concurrently under torchrun)."""
    import shutil

    from genvox_b200 import build
    assert build.is_fresh()                                   # the in-tree library matches the sources it ships with
    here = build._fingerprint()
    moved = tmp_path / "elsewhere"
    shutil.copytree(build.CSRC, moved / "csrc")
    shutil.copytree(build.INCLUDE, moved / "include")
    monkeypatch.setattr(build, "CSRC", str(moved / "csrc"))
    monkeypatch.setattr(build, "INCLUDE", str(moved / "include"))
    assert build._fingerprint() == here
    with open(moved / "csrc" / "gvx_common.cuh", "a") as fh:  # ... and it does depend on the contents
        fh.write("\n// touched\n")
    assert build._fingerprint() != here
