"""CPU: the C-ABI library loads and exports exactly the entry points include/mdb200.h declares; without a GPU the
product path fails loudly (no CPU fallback, no oracle on the product path)."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "mdb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mdb_[a-z_0-9]+)\s*\(", src)))


def test_header_matches_binding_and_library(md):
    hs = header_symbols()
    assert hs == sorted(md._capi.SYMBOLS)
    lib = md._capi.load()
    for name in hs:
        assert hasattr(lib, name), name
    out = subprocess.check_output(["nm", "-D", "--defined-only", md._capi.lib_path()]).decode()
    exported = sorted(set(re.findall(r" T (mdb_[a-z_0-9]+)", out)))
    assert exported == hs
    assert lib.mdb_version() == 100


def test_struct_layouts_match_header(md):
    import ctypes as C
    assert C.sizeof(md._capi.Config) == 4 + 4 + 8 + 72 + 8 + 64 + 8 + 4 + 4 + 8 + 4 + 4 + 4 + 4 + 8 + 8
    assert C.sizeof(md._capi.FireParams) == 8 + 6 * 8 + 8


def test_struct_offsets_match_a_c_compiler(md, tmp_path):
    """the ctypes mirrors of mdb_config / mdb_stats / mdb_fire_params have the size and field offsets gcc gives the header"""
    import ctypes as C
    structs = {"mdb_config": md._capi.Config, "mdb_stats": md._capi.Stats, "mdb_fire_params": md._capi.FireParams}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "mdb200.h"', 'int main(void) {']
    for cname, cls in structs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines += ['return 0;', '}']
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    got = dict(l.split() for l in subprocess.check_output([str(exe)]).decode().splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got["%s.%s" % (cname, fname)]) == getattr(cls, fname).offset, (cname, fname)


def test_product_path_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "moleculardynamics.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".inl")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "mdoracle" not in txt and "libmdoracle" not in txt and "oracle/" not in txt.replace("oracle/md_oracle.c", ""), f


def test_fails_loudly_without_gpu(md):
    try:
        import torch
        has = torch.cuda.is_available()
    except Exception:
        has = False
    if has:
        pytest.skip("GPU present")
    with pytest.raises(md.MdbError) as ei:
        md.Engine(3, 16, 10.0, 1.5, 0)
    assert ei.value.code == md._capi.ERR_NO_DEVICE
    p = md.Parameters(0.5, 16, 1e-3, md.PseudoHS())
    with pytest.raises(md.MdbError):
        md.initialize_state(p, None, random_init=True, write_init=False)


def test_hot_kernels_keep_their_register_budget(md):
    """compile-time guard (no GPU needed): the dominant kernel was tuned to 72 registers / 7 CTAs per SM with a small
    stack frame -- register spills to local memory cost it 10 % when they crept in (profiles/r01_session2.md)"""
    import shutil
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(tool):
        pytest.skip("cuobjdump not available")
    out = subprocess.check_output([tool, "--dump-resource-usage", md._capi.lib_path()]).decode()
    usage = {}
    name = None
    for line in out.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            name = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+)", line)
        if m and name:
            usage[name] = (int(m.group(1)), int(m.group(2)))
            name = None
    def find(*parts):
        hits = [v for k, v in usage.items() if all(p in k for p in parts)]
        assert hits, parts
        return hits
    # k_force_list<3, PotPHS, KICK2, SLAB=0, TRI=0>: fused NVE (2), plain second kick (1), forces only (0), fused Brownian (3)
    for kick, max_stack in ((2, 64), (1, 32), (0, 16), (3, 32)):
        for reg, stack in find("k_force_listILi3ENS_6PotPHSELi%dELb0ELb0" % kick):
            assert reg <= 72 and stack <= max_stack, (kick, reg, stack)
    for reg, stack in find("k_kick_driftILi3"):
        assert reg <= 64 and stack == 0
    for reg, stack in find("k_build_listILi3ELb0"):
        assert reg <= 64 and stack == 0
