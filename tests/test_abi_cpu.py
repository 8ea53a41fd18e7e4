"""No-GPU checks of the drop-in boundary: libdge.so loads and exports every symbol that
include/*.h declares; the config struct layouts agree; the product path refuses to run without a GPU."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    syms = []
    for h in ("dge.h", "dge_gnn.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        syms += re.findall(r"\b(dge_[a-z0-9_]+)\s*\(", src)
    return sorted(set(syms))


def test_library_exports_every_declared_symbol():
    from drl_graph_exploration_b200 import build_ext
    lib = ctypes.CDLL(build_ext.build())
    missing = [s for s in _declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    assert len(_declared_symbols()) >= 20


def test_config_struct_layout_matches_oracle():
    from drl_graph_exploration_b200.config import DgeConfigStruct, EnvConfig
    from oracle import oracle
    assert oracle.lib().orc_sizeof_config() == ctypes.sizeof(DgeConfigStruct)
    c = EnvConfig(map_size=60).to_struct()
    assert (c.map_min_x, c.map_max_x, c.env_max_y, c.num_landmarks) == (-50.0, 50.0, 30.0, 18)
    assert EnvConfig(map_size=100).rows == 70 and EnvConfig(map_size=20).rows == 30


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_product_path_fails_loudly_without_gpu():
    from drl_graph_exploration_b200.config import EnvConfig
    from drl_graph_exploration_b200.engine import DgeError, Engine
    from drl_graph_exploration_b200 import gnn
    with pytest.raises(DgeError):
        Engine(EnvConfig(map_size=20), 2)
    with pytest.raises(DgeError):
        gnn.GraphStructure(torch.zeros(2, 4, dtype=torch.long), None, 3)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "drl_graph_exploration_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f


def _header_prototypes():
    """name -> number of parameters, for every function include/*.h declares."""
    protos = {}
    for h in ("dge.h", "dge_gnn.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        src = re.sub(r"//[^\n]*", "", src)
        for name, args in re.findall(r"\b(dge_[a-z0-9_]+)\s*\(([^()]*)\)\s*;", src):
            args = args.strip()
            protos[name] = 0 if args in ("", "void") else len(args.split(","))
            kinds[name] = [] if args in ("", "void") else [_c_kind(a) for a in args.split(",")]
    return protos


kinds = {}


def _c_kind(decl: str) -> str:
    """'ptr' | 'i32' | 'i64' | 'f32' | 'f64' of one C parameter declaration."""
    d = decl.strip()
    if "*" in d or re.search(r"\bdge_handle\b", d):
        return "ptr"
    for pat, k in ((r"\b(u?int64_t|long long|size_t)\b", "i64"), (r"\b(u?int32_t|int|unsigned)\b", "i32"), (r"\bfloat\b", "f32"), (r"\bdouble\b", "f64")):
        if re.search(pat, d):
            return k
    raise AssertionError(f"unclassified parameter: {decl!r}")


def _ctypes_kind(t) -> str:
    if t is ctypes.c_void_p or t is ctypes.c_char_p or hasattr(t, "contents") or (isinstance(t, type) and issubclass(t, ctypes._Pointer)):
        return "ptr"
    return {ctypes.c_int: "i32", ctypes.c_int32: "i32", ctypes.c_uint32: "i32", ctypes.c_int64: "i64", ctypes.c_uint64: "i64",
            ctypes.c_float: "f32", ctypes.c_double: "f64"}[t]


def test_ctypes_argtypes_match_the_header_prototypes():
    """Every ``L.dge_*.argtypes = [...]`` in the package lists as many arguments as the prototype in include/*.h (a ctypes
    binding with one pointer too few or too many still 'works' until the kernel reads a garbage argument)."""
    protos = _header_prototypes()
    assert len(protos) >= 40 and protos["dge_create"] == 5 and protos["dge_last_error"] == 0
    from drl_graph_exploration_b200.config import DgeConfigStruct
    from drl_graph_exploration_b200.engine import _StateView
    vp = ctypes.c_void_p
    ns = {"ctypes": ctypes, "vp": vp, "_vp": vp, "DgeConfigStruct": DgeConfigStruct, "_StateView": _StateView}
    pkg = os.path.join(ROOT, "drl_graph_exploration_b200")
    seen, bad = set(), []
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if not f.endswith(".py"):
                continue
            src = open(os.path.join(dp, f)).read()
            for name, expr in re.findall(r"\.(dge_[a-z0-9_]+)\.argtypes\s*=\s*(\[[^\]]*\](?:\s*[+*]\s*(?:\[[^\]]*\]|\d+))*)\s*$", src, flags=re.M):
                types_ = eval(expr, dict(ns))          # noqa: S307 -- our own source, plain list arithmetic
                n = len(types_)
                seen.add(name)
                if name not in protos:
                    bad.append((f, name, "not declared in include/*.h"))
                elif protos[name] != n:
                    bad.append((f, name, f"argtypes has {n} entries, the prototype {protos[name]}"))
                elif [_ctypes_kind(t) for t in types_] != kinds[name]:      # pointer / 32-bit / 64-bit / float / double, position by position
                    bad.append((f, name, f"argtypes kinds {[_ctypes_kind(t) for t in types_]} != prototype {kinds[name]}"))
    assert not bad, bad
    unbound = sorted(set(protos) - seen - {"dge_last_error"})
    assert not unbound, f"no argtypes for {unbound}"


def _header_structs():
    """struct name -> [(field, kind)] in declaration order, for every ``typedef struct name { ... } name;`` of include/*.h."""
    out = {}
    for h in ("dge.h", "dge_gnn.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        src = re.sub(r"//[^\n]*", "", src)
        for name, body in re.findall(r"typedef\s+struct\s+(\w+)\s*\{(.*?)\}\s*\w+\s*;", src, flags=re.S):
            fields = []
            for decl in body.split(";"):
                decl = decl.strip()
                if not decl:
                    continue
                first, *rest = [d.strip() for d in decl.split(",")]
                base = re.sub(r"[\*\s]*\w+$", "", first).strip()             # the type shared by a `T a, *b, c;` declaration
                for d in [first] + rest:
                    fname = re.search(r"(\w+)$", d).group(1)
                    is_ptr = "*" in (d if d is not first else first[len(base):])
                    fields.append((fname, "ptr" if is_ptr else _c_kind(base + " x")))
            out[name] = fields
    return out


def test_ctypes_structures_mirror_the_header_structs():
    """Field names, order and kinds (pointer / 32-bit / 64-bit / double) of every ctypes.Structure the package hands to the
    library, against the struct of the same role in include/dge.h."""
    from drl_graph_exploration_b200.config import DgeConfigStruct
    from drl_graph_exploration_b200.engine import GraphOut, _StateView
    from drl_graph_exploration_b200.runner import GcnPolicy, GraphHostOut, GraphPacked, HostLoop
    hs = _header_structs()
    for cname, py in (("dge_config", DgeConfigStruct), ("dge_state_view", _StateView), ("dge_graph_out", GraphOut),
                      ("dge_graph_host_out", GraphHostOut), ("dge_graph_packed", GraphPacked), ("dge_gcn_policy", GcnPolicy),
                      ("dge_host_loop", HostLoop)):
        want = hs[cname]
        got = [(n, _ctypes_kind(t)) for n, t in py._fields_]
        assert got == want, (cname, [(a, b) for a, b in zip(got, want) if a != b], len(got), len(want))
    assert ctypes.sizeof(GraphPacked) == 14 * 8 + 4 * 4 and ctypes.sizeof(GraphHostOut) == 13 * 8


def test_call_sites_pass_as_many_arguments_as_the_prototypes_take():
    """Every ``<lib>.dge_*(...)`` call in the package, bench.py and the GPU tests' helpers passes the prototype's number of
    arguments (calls with *args are skipped: gnn.QForwardPlan's are covered by tests/test_q_plan_cpu.py)."""
    import ast
    protos = _header_prototypes()
    files = [os.path.join(dp, f) for dp, _, fs in os.walk(os.path.join(ROOT, "drl_graph_exploration_b200")) for f in fs if f.endswith(".py")]
    files += [os.path.join(ROOT, "bench.py")]
    bad, n_calls = [], 0
    for path in files:
        for node in ast.walk(ast.parse(open(path).read())):
            if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr in protos:
                if any(isinstance(a, ast.Starred) for a in node.args) or node.keywords:
                    continue
                n_calls += 1
                if len(node.args) != protos[node.func.attr]:
                    bad.append((os.path.basename(path), node.lineno, node.func.attr, len(node.args), protos[node.func.attr]))
    assert not bad, bad
    assert n_calls >= 50


def test_integration_doc_stub_lists_the_config_fields_of_the_header():
    """INTEGRATION.md shows the ctypes binding a reference maintainer would write; its dge_config field list is the header's."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = doc[doc.index("class dge_config(ctypes.Structure)"):doc.index("vp = ctypes.c_void_p")]
    names = re.findall(r'"(\w+)"', block)
    assert names == [n for n, _ in _header_structs()["dge_config"]]
