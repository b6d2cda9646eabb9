"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads without a GPU, exports every symbol that
``include/amp_b200.h`` declares, the ctypes table mirrors the header, and the product fails loudly without CUDA."""

from __future__ import annotations

import ctypes as C
import os
import re

import pytest
import torch

from conftest import ROOT, clip_path
from humanoid_amp_b200 import _lib

HEADER = os.path.join(ROOT, "include", "amp_b200.h")


def _declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = re.findall(r"AMP_API\s+([\w\s\*]+?)\s*\**\s*(amp_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S)
    out = {}
    for _ret, name, args in protos:
        args = args.strip()
        out[name] = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
    return out


def test_library_is_built_and_loads_without_gpu():
    assert os.path.exists(_lib.LIB_PATH), "run `python -m humanoid_amp_b200.build` (the driver calls __graft_entry__.build())"
    lib = _lib.load()
    assert lib.amp_b200_abi_version() == _lib.ABI_VERSION


def test_every_header_symbol_is_exported_and_typed():
    declared = _declared_functions()
    assert len(declared) >= 20
    lib = C.CDLL(_lib.LIB_PATH)
    for name, nargs in declared.items():
        assert hasattr(lib, name), f"{name} declared in amp_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} missing from the ctypes signature table"
        assert len(_lib.SIGNATURES[name][1]) == nargs, f"{name}: header has {nargs} args, ctypes table {len(_lib.SIGNATURES[name][1])}"
    assert set(_lib.SIGNATURES) == set(declared), "ctypes table and header disagree"


def test_desc_struct_layout_matches_header():
    # amp_lib_desc_t: 8 + 4*4 + 8 + 3*8 + 6*8 + 8 + 4 + 4 + 8 + 4 + 4 = 136 bytes, all pointers 8-byte aligned
    assert C.sizeof(_lib.LibDesc) == 136
    assert _lib.LibDesc.dt.offset == 24 and _lib.LibDesc.dof_positions.offset == 56
    assert _lib.LibDesc.dof_indexes.offset == 104 and _lib.LibDesc.key_body_indexes.offset == 120


def test_binary_carries_sm100a_tensor_core_and_tma_code():
    """cuobjdump the built library: the discriminator kernel must contain tcgen05 MMA (UTC*MMA), TMEM loads (LDTM) and
    TMA (UTMALDG) SASS for sm_100a -- the evidence B200_PROFILING.md asks for."""
    import shutil
    import subprocess

    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([exe, "-sass", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR"):
        assert mnemonic in sass, f"{mnemonic} not found in SASS"


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_fails_loudly_without_cuda():
    import humanoid_amp_b200 as amp

    with pytest.raises(amp.AmpB200Error):
        amp.MotionLoader(clip_path("G1_walk"), "cuda")
    with pytest.raises(amp.AmpB200Error):
        amp.MotionLoader(clip_path("G1_walk"), "cpu")  # there is no CPU path, by design
    with pytest.raises(amp.AmpB200Error):
        amp.AmpDiscriminator(166, device="cuda")
    lib = _lib.load()
    assert lib.amp_device_info(None, None, None) == _lib.AMP_ENODEV
    assert b"CUDA" in lib.amp_last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "humanoid_amp_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{fn} imports the oracle"


def test_argument_block_layouts_match_the_c_compiler(tmp_path):
    """``amp_lib_desc_t`` / ``amp_env_step_t``: size and every field offset of the ctypes mirrors against what gcc makes of
    ``include/amp_b200.h``."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    structs = {"amp_lib_desc_t": _lib.LibDesc, "amp_env_step_t": _lib.EnvStepArgs}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "amp_b200.h"', "int main(void) {"]
    for cname, cls in structs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run([gcc, "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = dict(line.split() for line in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, f"{cname}.{fname}"
