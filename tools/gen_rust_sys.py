#!/usr/bin/env python
"""Regenerate the `extern "C"` block of rust/cozo_gpu_sys.rs from include/cozo_gpu.h
(tests/test_abi_cpu.py::test_rust_ffi_matches_header checks the result against the header)."""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCALARS = {"int": "c_int", "int32_t": "i32", "uint32_t": "u32", "uint64_t": "u64", "int64_t": "i64", "float": "f32",
           "double": "f64", "uint8_t": "u8", "char": "c_char", "void": "c_void", "size_t": "usize"}
OPAQUE = {"cozo_gpu_hnsw_t": "CozoGpuHnsw", "cozo_gpu_graph_t": "CozoGpuGraph", "cozo_gpu_shards_t": "CozoGpuShards"}


def rust_type(c: str) -> str:
    c = c.replace("volatile", " ").strip()
    depth = c.count("*")
    base = c.replace("*", " ").split()
    const = base[0] == "const"
    if const:
        base = base[1:]
    name = base[0]
    t = SCALARS.get(name) or OPAQUE.get(name) or name
    if depth == 0:
        return t
    # the innermost pointer carries the C constness; outer levels are *mut (out parameters)
    out = ("*const " if const else "*mut ") + t
    for _ in range(depth - 1):
        out = "*mut " + out
    return out


def prototypes():
    hdr = open(os.path.join(ROOT, "include", "cozo_gpu.h")).read()
    body = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    for ret, name, args in re.findall(r"\n([A-Za-z_][A-Za-z_0-9 \*]*?)\s*\b(cozo_gpu_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", body, flags=re.S):
        args = " ".join(args.split())
        alist = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
        yield " ".join(ret.split()), name, alist


def main():
    lines = []
    for ret, name, args in prototypes():
        rargs = []
        for a in args:
            m = re.match(r"(.*?)([A-Za-z_][A-Za-z_0-9]*)$", a)
            rargs.append(f"{m.group(2)}: {rust_type(m.group(1))}")
        r = "" if ret == "void" else f" -> {rust_type(ret)}"
        lines.append(f"    pub fn {name}({', '.join(rargs)}){r};")
    p = os.path.join(ROOT, "rust", "cozo_gpu_sys.rs")
    src = open(p).read()
    head, rest = src.split('extern "C" {', 1)
    tail = rest.split("\n}\n", 1)[1] if "\n}\n" in rest else ""
    open(p, "w").write(head + 'extern "C" {\n' + "\n".join(lines) + "\n}\n" + tail)
    print(f"{len(lines)} entry points written to {p}")


if __name__ == "__main__":
    sys.exit(main())
