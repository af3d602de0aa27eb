"""CPU tests: libcozo_gpu.so loads and exports every symbol include/cozo_gpu.h declares
(no compute calls without a GPU), and fails loudly without a device."""
import ctypes
import os
import re

import pytest

from cozo_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from cozo_b200 import build
    build.build()
    return capi.load()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "cozo_gpu.h")).read()
    declared = set(re.findall(r"\b(cozo_gpu_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(capi.EXPORTS)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} is declared in cozo_gpu.h but not exported"


def test_struct_layouts_match_header(tmp_path):
    """sizeof/offsetof of the ctypes mirrors == what gcc computes from cozo_gpu.h"""
    import subprocess
    structs = {"CozoGpuHnswLevel": capi.HnswLevel, "CozoGpuHnswStageDesc": capi.HnswStageDesc,
               "CozoGpuSearchStats": capi.SearchStats, "CozoGpuHnswBuildDesc": capi.HnswBuildDesc}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "cozo_gpu.h"', 'int main(void){']
    for cname, cls in structs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines.append('return 0;}')
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, f"{cname}.{fname}"


def test_no_cpu_fallback_without_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert capi.device_count() == 0
    with pytest.raises(capi.CozoGpuError) as ei:
        capi.init(0)
    assert ei.value.code == capi.E_NODEV
    assert "no CPU fallback" in ei.value.msg


def test_product_never_imports_oracle():
    # the oracle is test infrastructure; nothing under cozo_b200/ may import, link or load it
    pat = re.compile(r"(from\s+oracle|import\s+oracle|oracle[/\\.]|libcozo_oracle|cozo_oracle)")
    for dirpath, _, files in os.walk(os.path.join(ROOT, "cozo_b200")):
        if os.path.basename(dirpath) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert not pat.search(txt), f"{f} references the oracle"


def _c_prototypes():
    hdr = open(os.path.join(ROOT, "include", "cozo_gpu.h")).read()
    body = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    out = {}
    for name, args in re.findall(r"\b(cozo_gpu_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", body, flags=re.S):
        args = " ".join(args.split())
        out[name] = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
    return out


def test_rust_ffi_matches_header():
    """rust/cozo_gpu_sys.rs (the binding a cozo maintainer compiles) declares exactly the header's
    entry points with the same arity, pointer depth and constness per argument."""
    src = open(os.path.join(ROOT, "rust", "cozo_gpu_sys.rs")).read()
    rust = {}
    for name, args in re.findall(r"pub fn (cozo_gpu_[a-z0-9_]+)\(([^)]*)\)", src):
        rust[name] = [a.strip() for a in args.split(",")] if args.strip() else []
    c = _c_prototypes()
    assert set(rust) == set(c) == set(capi.EXPORTS)
    for name, cargs in c.items():
        rargs = rust[name]
        assert len(rargs) == len(cargs), name
        for ca, ra in zip(cargs, rargs):
            cname = re.search(r"([A-Za-z_][A-Za-z_0-9]*)$", ca).group(1)
            rname, rty = [x.strip() for x in ra.split(":", 1)]
            assert cname == rname, (name, ca, ra)
            assert ca.count("*") == rty.count("*"), (name, ca, ra)
            if "*" in ca:
                assert ca.startswith("const ") == rty.startswith("*const "), (name, ca, ra)
    # repr(C) structs carry the same field names in the same order
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "cozo_gpu.h")).read(), flags=re.S)
    for sname in ("CozoGpuHnswLevel", "CozoGpuHnswStageDesc", "CozoGpuSearchStats", "CozoGpuHnswBuildDesc"):
        cbody = re.search(r"typedef struct \{([^}]*)\}\s*" + sname + ";", hdr, flags=re.S).group(1)
        cfields = re.findall(r"([A-Za-z_0-9]+)\s*;", cbody)
        rbody = re.search(r"pub struct " + sname + r" \{([^}]*)\}", src, flags=re.S).group(1)
        rfields = re.findall(r"pub ([A-Za-z_0-9]+):", rbody)
        assert cfields == rfields, sname
