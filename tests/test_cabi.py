"""The C-ABI library loads and exports every symbol include/vfengine.h declares; struct layouts of the
ctypes binding match the header; the product path fails loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vfengine.h")


@pytest.fixture(scope="module")
def lib():
    from visual_foresight_b200 import build, engine
    build.build()
    return engine.load_library()


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"VF_API\s+[\w\s\*]+?\b(vf_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from visual_foresight_b200 import engine
    decl = declared_symbols()
    assert len(decl) >= 25
    for name in decl:
        assert hasattr(lib, name), "libvfengine.so does not export %s" % name
    assert sorted(engine.exported_symbols()) == decl, "ctypes binding and header disagree"
    assert lib.vf_abi_version() == engine.VF_ABI_VERSION == 2


def test_struct_layout_matches_header(tmp_path):
    from visual_foresight_b200 import engine
    src = tmp_path / "sz.c"
    src.write_text('#include "%s"\n#include <stdio.h>\n#include <stddef.h>\nint main(){printf("%%zu %%zu %%zu %%zu %%zu\\n",'
                   'sizeof(vf_config),sizeof(vf_cem_params),sizeof(vf_tensor),offsetof(vf_cem_params,seed),'
                   'offsetof(vf_config,max_samples));return 0;}\n' % HEADER)
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", str(src), "-o", str(exe)])          # header must also be valid C
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [ctypes.sizeof(engine.VfConfig), ctypes.sizeof(engine.VfCemParams), ctypes.sizeof(engine.VfTensor),
            engine.VfCemParams.seed.offset, engine.VfConfig.max_samples.offset]
    assert got == want


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from visual_foresight_b200 import engine, spec
    with pytest.raises(engine.EngineUnavailable):
        engine.Engine(spec.spec_64(), 4)
    from visual_foresight_b200.cem_controller import PixelCostController
    ag = {"adim": 4, "sdim": 4, "image_height": 64, "image_width": 64, "gpu_id": 0}
    with pytest.raises(engine.EngineUnavailable):
        PixelCostController(ag, {"rejection_sampling": False, "verbose": False}, 0, 1)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "visual_foresight_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), "%s imports oracle" % f
                assert "/root/reference" not in txt or f.endswith(".py") and "reference" in txt
