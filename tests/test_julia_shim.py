"""The Julia `ccall` shim cannot be executed here (no Julia in the image).  This is the next best thing: every
`ccall` in healpixmpi.jl_b200/julia/HealpixMPIB200.jl is parsed and checked against the prototype of the same symbol
in include/hpsht_b200.h (existence, return type, number and kind of arguments, as many values as types), and the
file's block structure (module / function / struct / if / for ... end) must balance.  Binding drift -- a symbol
renamed in the header, an argument added on one side only -- fails here instead of at a maintainer's first call."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JL = os.path.join(ROOT, "healpixmpi.jl_b200", "julia", "HealpixMPIB200.jl")
HDR = os.path.join(ROOT, "include", "hpsht_b200.h")


def _strip_julia(src):
    """comments and string literals out (no nested / triple-quoted strings in the shim)"""
    out = []
    for line in src.splitlines():
        buf, i, in_str = [], 0, False
        while i < len(line):
            ch = line[i]
            if in_str:
                if ch == "\\":
                    i += 2
                    continue
                if ch == '"':
                    in_str = False
                i += 1
                continue
            if ch == '"':
                in_str = True
                buf.append('""')
                i += 1
                continue
            if ch == "#":
                break
            buf.append(ch)
            i += 1
        out.append("".join(buf))
    return "\n".join(out)


def _split_top(s):
    """split at top-level commas"""
    parts, depth, cur = [], 0, []
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append("".join(cur).strip())
            cur = []
        else:
            cur.append(ch)
    tail = "".join(cur).strip()
    if tail:
        parts.append(tail)
    return parts


def _ccalls(src):
    """[(symbol, return type, [argument types], [argument values])]"""
    calls = []
    for m in re.finditer(r"\bccall\(", src):
        i, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
        parts = _split_top(src[m.end():i - 1])
        sym = re.match(r"\(\s*:(\w+)\s*,\s*lib\s*\)", parts[0])
        assert sym, parts[0]
        types = parts[2].strip()
        assert types.startswith("(") and types.endswith(")"), types
        tlist = _split_top(types[1:-1])
        calls.append((sym.group(1), parts[1].strip(), tlist, parts[3:]))
    return calls


def _prototypes(hdr):
    """{symbol: (return type, [parameter types])} from the C header"""
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    hdr = re.sub(r"//[^\n]*", " ", hdr)
    hdr = "\n".join(l for l in hdr.splitlines() if not l.lstrip().startswith("#") and 'extern "C"' not in l)
    protos = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(hpsht_\w+)\s*\(([^;{]*?)\)\s*;", hdr):
        ret, name, params = m.group(1).strip().split(";")[-1].split("}")[-1].strip(), m.group(2), m.group(3).strip()
        plist = [] if params in ("", "void") else [p.strip() for p in params.split(",")]
        ptypes = []
        for p in plist:
            p = re.sub(r"\bconst\b", "", p)
            p = re.sub(r"\s+", " ", p).strip()
            arr = re.match(r"(.*?)\s*(\w+)\s*\[[^\]]*\]$", p)   # `unsigned char id[N]` is `unsigned char*`
            if arr:
                p = arr.group(1).strip() + "* " + arr.group(2)
            t = re.match(r"(.*?)(\w+)?$", p)  # drop the parameter name
            base = t.group(1).strip() if t.group(2) and t.group(1).strip() else p
            ptypes.append(base.replace(" *", "*").replace("* ", "*"))
        protos[name] = (re.sub(r"\bconst\b", "", ret).strip().replace(" *", "*"), ptypes)
    return protos


# which Julia ccall types may stand for which C parameter type
COMPAT = {
    "int": {"Cint"},
    "int64_t": {"Int64"},
    "double": {"Float64", "Cdouble"},
    "double*": {"Ptr{Float64}", "Ptr{ComplexF64}", "Ref{Float64}"},
    "int64_t*": {"Ptr{Int64}", "Ref{Int64}"},
    "int*": {"Ptr{Cint}", "Ref{Cint}"},
    "unsigned char*": {"Ptr{UInt8}"},
    "void*": {"Ptr{Cvoid}"},
    "hpsht_plan*": {"Ptr{Cvoid}"},
    "hpsht_plan**": {"Ref{Ptr{Cvoid}}", "Ptr{Ptr{Cvoid}}"},
    "char*": {"Cstring", "Ptr{UInt8}"},
}


def test_every_ccall_matches_its_prototype():
    src = _strip_julia(open(JL).read())
    protos = _prototypes(open(HDR).read())
    calls = _ccalls(src)
    assert len(calls) >= 20
    for sym, ret, types, values in calls:
        assert sym in protos, f"{sym} is not declared in include/hpsht_b200.h"
        cret, cparams = protos[sym]
        assert ret in COMPAT[cret], (sym, ret, cret)
        assert len(types) == len(cparams), (sym, types, cparams)
        assert len(values) == len(types), (sym, "values handed over", values, "types", types)
        for jt, ct in zip(types, cparams):
            assert ct in COMPAT, (sym, ct)
            assert jt in COMPAT[ct], (sym, jt, "cannot stand for", ct)


def test_fused_entry_points_are_bound():
    """the symbols the drop-in boundary consists of are all reachable from Julia"""
    src = _strip_julia(open(JL).read())
    bound = {c[0] for c in _ccalls(src)}
    for sym in ("hpsht_alm2leg", "hpsht_leg2alm", "hpsht_leg2map", "hpsht_map2leg", "hpsht_plan_create",
                "hpsht_plan_destroy", "hpsht_plan_set_geometry", "hpsht_alm2map", "hpsht_adjoint_alm2map",
                "hpsht_alm2map_async", "hpsht_adjoint_alm2map_async", "hpsht_plan_synchronize",
                "hpsht_nccl_unique_id", "hpsht_plan_sizes", "hpsht_plan_alm_info", "hpsht_plan_map_info",
                "hpsht_last_error"):
        assert sym in bound, sym


@pytest.mark.parametrize("path", [JL, os.path.join(ROOT, "tests", "golden", "make_reference_golden.jl"),
                                  os.path.join(ROOT, "bench", "reference_cpu.jl")])
def test_blocks_balance_and_names_are_unique(path):
    src = _strip_julia(open(path).read())
    openers = {"module", "function", "struct", "if", "for", "while", "let", "begin", "try", "quote", "macro", "do"}
    depth, stack = 0, []
    for ln, line in enumerate(src.splitlines(), 1):
        toks = re.findall(r"[A-Za-z_]\w*!?", line)
        # statement-leading openers only: `x for x in ...` comprehensions and `a ? b : c` open nothing
        stmts = [s.strip() for s in line.split(";")]
        for st in stmts:
            first = re.match(r"(mutable\s+struct|[A-Za-z_]\w*)", st)
            if first:
                word = first.group(1).split()[-1]
                if word in openers:
                    stack.append((word, ln))
            if re.search(r"\bdo\b(\s+\w+)?\s*$", st) or re.search(r"=\s*begin\b", st):
                stack.append(("do/begin", ln))
        for t in toks:
            if t == "end":
                # `end` inside an index expression (a[end]) does not occur in the shim
                assert stack, f"line {ln}: `end` without an opener"
                stack.pop()
    assert not stack, f"unclosed blocks: {stack}"
    del depth
    # a definition pasted twice (round-1 defect: a copy of the fused helpers inside `module Sht`) is caught here
    names = re.findall(r"^\s*function\s+([\w\.!]+)\s*\(", src, flags=re.M)
    names = [n for n in names if n != "Plan"]
    assert len(names) == len(set(names)), sorted(n for n in names if names.count(n) > 1)


def test_ctypes_table_matches_the_prototypes():
    """Same check for the Python binding: argument count, scalar width and pointer-ness of every entry of
    healpixmpi.jl_b200/_lib.py::SIGNATURES against the header (a c_int where the C side takes int64_t would pass
    garbage in the upper half on this ABI only by luck)."""
    import ctypes as C
    import importlib

    lib = importlib.import_module("healpixmpi_jl_b200")._lib
    protos = _prototypes(open(HDR).read())
    scalar = {"int": C.c_int, "int64_t": C.c_int64, "size_t": C.c_size_t, "double": C.c_double}
    for name, (restype, argtypes) in lib.SIGNATURES.items():
        assert name in protos, name
        cret, cparams = protos[name]
        assert len(argtypes) == len(cparams), (name, argtypes, cparams)
        assert restype is (C.c_char_p if cret == "char*" else scalar[cret]), (name, restype, cret)
        for at, ct in zip(argtypes, cparams):
            if ct.endswith("*"):
                is_ptr = at is C.c_void_p or at is C.c_char_p or hasattr(at, "_type_") and isinstance(at._type_, type)
                assert is_ptr, (name, at, ct)
                if at not in (C.c_void_p, C.c_char_p):   # a typed POINTER must point at the right scalar
                    base = ct[:-1].strip()
                    want = C.c_void_p if base.endswith("*") or base in ("void", "hpsht_plan") else scalar.get(base)
                    assert want is not None and at._type_ is want, (name, at, ct)
            else:
                assert at is scalar[ct], (name, at, ct)
    assert set(protos) == set(lib.SIGNATURES)
