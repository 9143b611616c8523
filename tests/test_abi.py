"""CPU suite: the C-ABI library loads, exports every symbol include/d3fk.h declares, the Python struct mirror
matches, and the product fails loudly (no fallback) when there is no sm_100 device."""
import ctypes
import os
import re

import pytest
import torch

import denoising_diffusion_deep_fake_b200 as d3
from denoising_diffusion_deep_fake_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAS_GPU = torch.cuda.is_available()


def header_functions():
    text = open(os.path.join(ROOT, "include", "d3fk.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(d3fk_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"libd3fk.so does not export {n}"
    assert set(_lib.EXPORTS) <= set(names)


def test_struct_mirror_matches_library():
    lib = _lib.load()
    assert lib.d3fk_sizeof_op() == ctypes.sizeof(_lib.Op)
    assert lib.d3fk_version() >= 1
    op = _lib.make_op(_lib.OP_CONV, B=2, Hi=8, Wi=8, Cout=16, src0=1234)
    assert op.kind == _lib.OP_CONV and _lib.op_params(op).Cout == 16 and _lib.op_params(op).src0 == 1234
    with pytest.raises(KeyError):
        _lib.make_op(_lib.OP_CONV, not_a_field=1)


@pytest.mark.skipif(HAS_GPU, reason="checks the no-GPU failure path")
def test_no_device_is_an_error_not_a_fallback():
    lib = _lib.load()
    assert lib.d3fk_init(0) == -2                                   # D3FK_ERR_ARCH
    assert b"no CPU path" in lib.d3fk_last_error() or b"sm_100" in lib.d3fk_last_error()
    arr = (_lib.Op * 1)(_lib.make_op(_lib.OP_INC, p0=0, n=1))
    assert lib.d3fk_run(arr, 1, None) == -2                          # refuses to run uninitialised
    with pytest.raises(d3.D3fkError):
        d3.Unet(precision="fp32")(torch.randn(1, 3, 32, 32))
    with pytest.raises(d3.D3fkError):
        d3.q_sample(torch.randn(1, 3, 32, 32), 5.0)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "denoising_diffusion_deep_fake_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_op_kinds_match_the_header_enum():
    """Every D3FK_OP_* value of include/d3fk.h equals the OP_* constant of the Python mirror (and vice versa)."""
    text = open(os.path.join(ROOT, "include", "d3fk.h")).read()
    body = text[text.index("enum d3fk_op_kind"):]
    body = re.sub(r"/\*.*?\*/", "", body[:body.index("};")], flags=re.S)
    header = {m.group(1): int(m.group(2)) for m in re.finditer(r"D3FK_OP_([A-Z0-9_]+)\s*=\s*(\d+)", body)}
    mirror = {k[3:]: v for k, v in vars(_lib).items() if k.startswith("OP_") and isinstance(v, int)}
    assert header == mirror, (sorted(set(header.items()) ^ set(mirror.items())))
    assert len(set(header.values())) == len(header)          # no two kinds share a value


def test_space_to_depth_stem_is_chosen_where_its_tile_is_a_box():
    """plan.UnetPlan._stem_box_ok mirrors tma_box() of csrc/conv_tc.cu for the stem's output grid."""
    from denoising_diffusion_deep_fake_b200.plan import UnetPlan
    ok = UnetPlan._stem_box_ok
    assert ok(32, 32) and ok(64, 64) and ok(128, 128) and ok(16, 16) and ok(32, 16) and ok(16, 256)
    assert not ok(48, 48) and not ok(80, 80)                 # 96 x 96 / 160 x 160 inputs keep the gather-form stem


def test_library_is_blackwell_tensor_core_code():
    """The shipped libd3fk.so is tcgen05 / TMA code, not a recompiled mma.sync port: every convolution / weight-gradient kernel
    carries UTCHMMA (tcgen05.mma) and LDTM (tcgen05.ld); the TMA-fed ones carry UTMALDG and no LDGSTS; MMA / TMA issue sits in
    elect.sync regions (no R2UR.BROADCAST waterfall loops); the slab weight gradient reduces with 16-byte
    REDG (item 8e); the only warp-level HMMA user is the N = 3 head (item 8h).  Reads SASS with cuobjdump: no GPU needed."""
    import shutil, subprocess, sys
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "sass_opcodes.py"), _lib.LIB_PATH], capture_output=True, text=True,
                         check=True).stdout.splitlines()
    cols = out[0].split()[4:]            # after "kernel (mangled, d3fk:: stripped)"
    rows = {}
    for line in out[1:]:
        parts = line.split()
        rows[parts[0]] = dict(zip(cols, map(int, parts[1:])))
    fam = {k: v for k, v in rows.items() if any(n in k for n in ("conv_tc_kernel", "conv_slab_kernel", "wgrad_tc_kernel", "wgrad_slab_kernel"))}
    assert len(fam) >= 20
    for k, v in fam.items():
        assert v["UTCHMMA"] > 0 and v["LDTM"] > 0 and v["UTCBAR"] > 0, k
        assert v["R2UR.BROADCAST"] == 0 and v["HMMA"] == 0, k
    slab = [v for k, v in fam.items() if "slab" in k]
    assert all(v["UTMALDG"] > 0 and v["LDGSTS"] == 0 for v in slab)
    tma_wgrad = [v for k, v in fam.items() if "wgrad_tc_kernel" in k and "ELb1E" in k]
    assert tma_wgrad and all(v["UTMALDG"] > 0 and v["LDGSTS"] == 0 for v in tma_wgrad)
    assert all(v["RED.E.ADD.F32x4"] > 0 for k, v in fam.items() if "wgrad_slab" in k)
    head = [v for k, v in rows.items() if "head_conv_mma" in k]
    assert head and all(v["HMMA"] == 9 and v["LDSM"] == 9 for v in head)
