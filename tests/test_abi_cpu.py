"""CPU-only checks of the C-ABI boundary: the shared library loads without a GPU, exports every symbol
include/pcseg_b200.h declares, and its host-side layout queries match the reference's state_dict."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pcseg_b200.h")
LIB = os.path.join(ROOT, "point-cloud-cnn-segmentation_b200", "lib", "libpcseg_b200.so")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        import __graft_entry__
        __graft_entry__.build()
    return ctypes.CDLL(LIB)


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pcseg_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    names = declared_symbols()
    assert len(names) >= 18
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_binding_table_matches_header():
    import pcseg_b200
    from pcseg_b200._lib import EXPORTS
    assert sorted(EXPORTS) == declared_symbols()


def test_layout_matches_reference_state_dict(lib):
    from oracle import pointnet_oracle as orc
    lib.pcseg_param_count.restype = ctypes.c_longlong
    lib.pcseg_param_offset.restype = ctypes.c_longlong
    lib.pcseg_param_numel.restype = ctypes.c_longlong
    lib.pcseg_bn_buffer_count.restype = ctypes.c_longlong
    for C in (3, 5):
        spec = [(n, s) for n, s, dt in orc.state_dict_spec(C) if dt == np.float32 and "running" not in n]
        assert len(spec) == 38
        off = 0
        for t, (name, shape) in enumerate(spec):
            assert lib.pcseg_param_offset(C, t) == off, name
            assert lib.pcseg_param_numel(C, t) == int(np.prod(shape)), name
            off += int(np.prod(shape))
        assert lib.pcseg_param_count(C) == off
    assert lib.pcseg_param_count(5) == 1927621
    assert lib.pcseg_bn_buffer_count() == 6528


def test_error_paths_without_gpu(lib):
    lib.pcseg_last_error.restype = ctypes.c_char_p
    lib.pcseg_workspace_bytes.restype = ctypes.c_longlong
    h = ctypes.c_void_p()
    assert lib.pcseg_create(ctypes.byref(h), 99) != 0
    assert b"num_classes" in lib.pcseg_last_error()
    assert lib.pcseg_create(ctypes.byref(h), 5) == 0
    assert lib.pcseg_forward_eval(h, None, None, None, None) != 0     # not bound
    assert lib.pcseg_workspace_bytes(8, 16384, 5, 1) > lib.pcseg_workspace_bytes(8, 16384, 5, 0) > 0
    assert lib.pcseg_workspace_bytes(0, 16, 5, 1) == -1
    lib.pcseg_destroy(h)


def test_module_is_a_state_dict_drop_in():
    """Same 65 state_dict entries / shapes / dtypes / order as the reference module (pcs.py:70-94), identical default
    initialisation under the same torch seed, and strict loading of a reference-layout checkpoint (pcs.py:373-382,
    401-430) with and without the DataParallel `module.` prefix."""
    import io
    import torch
    import pcseg_b200
    from oracle import pointnet_oracle as orc
    C = 5
    m = pcseg_b200.PointNetSegmentation(C)
    sd = m.state_dict()
    spec = orc.state_dict_spec(C)
    assert list(sd.keys()) == [n for n, _, _ in spec]
    for n, shape, dt in spec:
        assert tuple(sd[n].shape) == tuple(shape), n
        assert sd[n].dtype == (torch.int64 if dt == np.int64 else torch.float32), n
    # default init parity with torch.nn (what the reference constructor does)
    torch.manual_seed(3)
    a = pcseg_b200.PointNetSegmentation(C)
    torch.manual_seed(3)
    ref_conv = torch.nn.Conv1d(4, 64, 1)
    assert torch.equal(a.conv1.weight, ref_conv.weight) and torch.equal(a.conv1.bias, ref_conv.bias)
    # checkpoint round trip in the reference's dict layout
    synth = {k: torch.from_numpy(np.asarray(v)) for k, v in orc.synth_state(C, 1).items()}
    for prefix in ("", "module."):
        ckpt = {"epoch": 3, "model_state_dict": {prefix + k: v for k, v in synth.items()}, "optimizer_state_dict": {},
                "train_loss": 1.0, "val_loss": 1.1, "f1_class2": 0.5, "f1_per_class": [0.1] * C, "num_classes": C}
        buf = io.BytesIO()
        torch.save(ckpt, buf)
        buf.seek(0)
        model, loaded = pcseg_b200.load_checkpoint(buf)
        assert loaded["num_classes"] == C
        for k, v in synth.items():
            assert torch.equal(model.state_dict()[k], v), k
    with pytest.raises(NotImplementedError):
        pcseg_b200.PointNetSegmentation(C, input_dim=3)


def test_host_side_helpers_without_gpu():
    """host logic that needs no device: per-cloud lengths validation, capacity buckets of ragged calls, mask -> lengths,
    F1 scores from a confusion matrix (pcs.py:341-343) against sklearn"""
    import numpy as np
    import pytest
    import torch
    import pcseg_b200
    from pcseg_b200.engine import host_lengths, ragged_capacity
    from sklearn.metrics import confusion_matrix, f1_score

    assert host_lengths(None, 3, 10) is None
    assert host_lengths([10, 10, 10], 3, 10) is None                     # nothing padded -> dense path
    arr = host_lengths(torch.tensor([10, 0, 7]), 3, 10)
    assert list(arr) == [10, 0, 7]
    assert list(host_lengths(np.array([1, 2, 3]), 3, 10)) == [1, 2, 3]
    for bad in ([1, 2], [1, 2, 11], [-1, 2, 3]):
        with pytest.raises(ValueError):
            host_lengths(bad, 3, 10)
    assert ragged_capacity(1) == 4096 and ragged_capacity(4096) == 4096 and ragged_capacity(4097) == 8192

    masks = torch.zeros(3, 8, dtype=torch.bool)
    masks[0, :8] = True
    masks[1, :3] = True
    assert pcseg_b200.lengths_from_masks(masks) == [8, 3, 0]

    rng = np.random.default_rng(0)
    C = 6
    yt = rng.integers(0, C - 1, 500)          # class C-1 never occurs as a label ...
    yp = rng.integers(0, C, 500)              # ... but is predicted
    conf = torch.from_numpy(confusion_matrix(yt, yp, labels=list(range(C))))
    f1, macro, weighted = pcseg_b200.f1_scores(conf)
    np.testing.assert_allclose(f1.numpy(), f1_score(yt, yp, average=None, labels=list(range(C))), atol=1e-12)
    assert abs(macro.item() - f1_score(yt, yp, average="macro")) < 1e-12
    assert abs(weighted.item() - f1_score(yt, yp, average="weighted")) < 1e-12


def test_ragged_plan_invariants(lib):
    """pcseg_ragged_plan (pure host code, the layout every *_ragged call runs on): clouds start at multiples of 128 rows,
    hold their real rows plus ONE representative pad row when they are padded, tiles map to exactly one cloud, and the
    BN-backward strips tile the packed rows exactly once without crossing a cloud."""
    import numpy as np
    lib.pcseg_ragged_plan.restype = ctypes.c_longlong
    lib.pcseg_ragged_plan.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
                                      ctypes.c_longlong, ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(ctypes.c_int)]
    rng = np.random.default_rng(0)
    cases = [(3, 1000, [1000, 517, 1]), (2, 256, [256, 0]), (1, 100, [100]), (5, 777, [1, 2, 3, 776, 777]),
             (8, 16384, [16384] + rng.integers(1, 16384, 7).tolist()), (64, 65536, rng.integers(0, 65537, 64).tolist())]
    for B, N, lengths in cases:
        arr = (ctypes.c_int * B)(*lengths)
        rows, strips = ctypes.c_longlong(), ctypes.c_int()
        n = lib.pcseg_ragged_plan(B, N, arr, None, 0, ctypes.byref(rows), ctypes.byref(strips))
        assert n > 0
        meta = (ctypes.c_int * n)()
        assert lib.pcseg_ragged_plan(B, N, arr, meta, n, None, None) == n
        m = np.frombuffer(meta, dtype=np.int32)
        ln, off = m[:B], m[B:2 * B + 1]
        T = rows.value // 128
        tile_cloud = m[2 * B + 1:2 * B + 1 + T]
        st = m[2 * B + 1 + T:].reshape(-1, 4)
        assert list(ln) == lengths and off[0] == 0 and off[B] == rows.value
        assert rows.value <= B * ((N + 127) // 128 * 128)
        for b in range(B):
            alloc = off[b + 1] - off[b]
            need = lengths[b] + (1 if lengths[b] < N else 0)
            assert off[b] % 128 == 0 and alloc % 128 == 0 and need <= alloc < need + 128 and alloc >= 128
            assert (tile_cloud[off[b] // 128:off[b + 1] // 128] == b).all()
        assert len(st) == strips.value
        covered = np.zeros(rows.value, dtype=np.int32)
        for cl, r0, r1, base in st:
            assert off[cl] <= r0 < r1 <= off[cl + 1] and base == off[cl] and r0 % 128 == 0
            covered[r0:r1] += 1
        assert (covered == 1).all()
    # errors: length out of range, meta buffer too small
    bad = (ctypes.c_int * 2)(5, 11)
    assert lib.pcseg_ragged_plan(2, 10, bad, None, 0, None, None) == -1
    ok = (ctypes.c_int * 2)(5, 10)
    small = (ctypes.c_int * 4)()
    assert lib.pcseg_ragged_plan(2, 10, ok, small, 4, None, None) == -1
    assert b"too small" in lib.pcseg_last_error()


def test_header_is_plain_c():
    """the boundary is a C ABI: include/pcseg_b200.h must compile as C99 (and as C++) on its own, and a C host that calls
    the layout queries must link against the shared library"""
    import shutil
    import subprocess
    import tempfile
    if shutil.which("gcc") is None:
        import pytest
        pytest.skip("no gcc")
    hdr = os.path.join(ROOT, "include", "pcseg_b200.h")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr], check=True)
    subprocess.run(["g++", "-std=c++11", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++", hdr], check=True)
    src = r'''
#include <stdio.h>
#include "pcseg_b200.h"
int main(void) {
    int lengths[3] = {1000, 517, 1};
    long long rows = 0; int strips = 0;
    long long n = pcseg_ragged_plan(3, 1000, lengths, NULL, 0, &rows, &strips);
    printf("%lld %lld %lld %d\n", pcseg_param_count(5), pcseg_bn_buffer_count(), rows, strips > 0 && n > 0);
    return 0;
}
'''
    libdir = os.path.join(ROOT, "point-cloud-cnn-segmentation_b200", "lib")
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "host.c")
        exe = os.path.join(td, "host")
        with open(c, "w") as f:
            f.write(src)
        subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), c, "-o", exe, "-L", libdir, "-lpcseg_b200",
                        "-Wl,-rpath," + libdir], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()
    assert out[0] == "1927621" and out[1] == "6528" and int(out[2]) == 1024 + 640 + 128 and out[3] == "1"
