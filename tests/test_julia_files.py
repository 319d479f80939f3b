"""Static checks of the Julia side of the drop-in (julia/*.jl).  Julia is not installable in the build image, so the
files cannot be executed here; what can be pinned is everything they must agree on with the C side: the generated
constant tables, the `ccall` struct layouts against include/smm_b200.h, the objective ids, and that every symbol they
bind is exported by libsmm_b200.so."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
J = os.path.join(ROOT, "julia")


def _read(*p):
    with open(os.path.join(*p)) as f:
        return f.read()


def test_julia_tables_are_generated_from_the_header():
    from tools import gen_julia_tables as g
    assert g.main() == _read(J, "smm_stream_tables.jl"), "run: python tools/gen_julia_tables.py > julia/smm_stream_tables.jl"
    mac = g.macros(_read(ROOT, "include", "smm_stream_tables.h"))
    txt = re.sub(r"#.*", "", _read(J, "smm_stream_tables.jl"))

    def arr(name):
        body = txt.split(f"const {name} = ", 1)[1].split("]", 1)[0].split("[", 1)[1]
        return [t.strip() for t in body.split(",") if t.strip()]
    for name, n in (("SMM_SIN_COEFS", int(mac["SMM_SIN_DEG"]) + 1), ("SMM_COS_COEFS", int(mac["SMM_COS_DEG"]) + 1),
                    ("SMM_LOGQ_COEFS", int(mac["SMM_LOGQ_DEG"]) + 1), ("SMM_EXP_COEFS", int(mac["SMM_EXP_DEG"]) + 1),
                    ("SMM_ZIG_F", int(mac["SMM_ZIG_LAYERS"]) + 1), ("SMM_LOG_INV", 1 << int(mac["SMM_LOG_BITS"])),
                    ("SMM_LOG_NLNC", 1 << int(mac["SMM_LOG_BITS"]))):
        vals = arr(name)
        assert len(vals) == n, name
        assert all(float.fromhex(v) == float.fromhex(v) for v in vals)           # every literal is a C99 hex float
    z = arr("SMM_ZIG_TABLE")
    hz = [int(t.lower().rstrip("ul"), 16) for t in g.items(mac["SMM_ZIG_TABLE"])]
    assert [int(t, 16) for t in z] == hz and len(z) == int(mac["SMM_ZIG_LAYERS"])
    assert all(len(t) == 18 for t in z)                                          # 16 hex digits: a UInt64 literal in Julia
    hl = [tuple(float.fromhex(x.strip()) for x in e.strip("{} ").split(",")) for e in g.items(mac["SMM_LOG_TABLE"])]
    assert [float.fromhex(v) for v in arr("SMM_LOG_INV")] == [a for a, _ in hl]
    assert [float.fromhex(v) for v in arr("SMM_LOG_NLNC")] == [b for _, b in hl]


C2J = {"int32_t": "Int32", "uint64_t": "UInt64", "double": "Cdouble", "const double *": "Ptr{Cdouble}", "double *": "Ptr{Cdouble}",
       "uint8_t *": "Ptr{UInt8}", "int32_t *": "Ptr{Int32}"}


def _c_struct_fields(header: str, name: str):
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), header, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    out = []
    for decl in body.split(";"):
        decl = " ".join(decl.split())
        if not decl:
            continue
        m = re.match(r"(.*?)(\w+)(\[\w+\])?$", decl)
        ctype, field, arr = m.group(1).strip(), m.group(2), m.group(3)
        out.append((field, "NTuple{128,UInt8}" if arr else C2J[ctype]))
    return out


def _julia_struct_fields(src: str, name: str):
    body = re.search(r"^struct %s\n(.*?)^end" % name, src, flags=re.S | re.M).group(1)
    return [tuple(x.strip() for x in ln.split("::")) for ln in body.strip().splitlines()]


def test_ccall_structs_mirror_the_c_header():
    h = _read(ROOT, "include", "smm_b200.h")
    jl = _read(J, "AlgoBGPB200.jl")
    assert _julia_struct_fields(jl, "SmmBgpConfig") == _c_struct_fields(h, "smm_bgp_config")
    assert _julia_struct_fields(jl, "SmmTraceView") == _c_struct_fields(h, "smm_trace_view")
    assert "const SMM_ABI_VERSION = Int32(%s)" % re.search(r"#define SMM_ABI_VERSION (\d+)", h).group(1) in jl
    assert "const SMM_E_UNSUPPORTED_SHAPE = %s" % re.search(r"#define SMM_E_UNSUPPORTED_SHAPE \((-\d+)\)", h).group(1) in jl


def test_objective_ids_and_bound_symbols():
    h = _read(ROOT, "include", "smm_b200.h")
    jl = _read(J, "AlgoBGPB200.jl")
    ids = {k: int(v) for k, v in re.findall(r"#define (SMM_OBJ_\w+) (\d+)", h)}
    table = dict(re.findall(r"(\w+) => (\d+)", jl.split("const SMM_OBJ = ", 1)[1].split("\n\n", 1)[0]))
    assert table == {"objfunc_norm": str(ids["SMM_OBJ_NORM"]), "objfunc_norm_b200": str(ids["SMM_OBJ_NORM"]),
                     "objfunc_norm_slow": str(ids["SMM_OBJ_NORM_SLOW"]), "objfunc_norm_mv": str(ids["SMM_OBJ_NORM_MV"]),
                     "objfunc_panel": str(ids["SMM_OBJ_PANEL"]), "Testobj_fails": str(ids["SMM_OBJ_FAILS"])}
    from smm_jl_b200 import _lib
    bound = set(re.findall(r"ccall\(\(:(\w+), LIBSMM_B200\)", jl))
    assert bound and bound <= set(_lib.EXPORTS), bound - set(_lib.EXPORTS)
    for sym in bound:                       # ... and declared in the header
        assert re.search(r"\b%s\(" % sym, h), sym
    # the objectives the table names exist as Julia functions on the shared streams
    obj = _read(J, "ObjB200.jl")
    for f in ("objfunc_norm_b200", "objfunc_norm_mv", "objfunc_panel"):
        assert f"function {f}(ev::Eval; kw...)" in obj


def test_stream_port_covers_the_header():
    """every stream function of include/smm_stream.h has its Julia twin (names without the smm_ prefix)"""
    s = _read(J, "SMMStreams.jl")
    for f in ("philox4x32_10", "u01", "u01_open", "u32_to_double", "neglog01", "normal_pair", "exp_neg", "zig_select",
              "zig_slow", "zig_normal", "zig_triple", "sim_block", "sim_normals", "prop_normal", "acc_uniform",
              "pair_unrank", "pair_sample"):
        assert re.search(r"^(@inline )?(function )?%s\(" % f, s, flags=re.M), f
    h = _read(ROOT, "include", "smm_stream.h")
    for name, val in re.findall(r"#define (SMM_PHILOX_\w+|SMM_ZIG_TAG|SMM_ZIG_KEY\d) (0x[0-9A-Fa-f]+)u", h):
        assert re.search(r"const \w+ = %s\b" % val, s, flags=re.I), name
