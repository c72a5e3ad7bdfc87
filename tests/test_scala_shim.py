"""The reference-side binding shown in INTEGRATION.md (scala/.../GingrCudaNative.scala, Panama FFM) cannot be compiled
here (no JVM), so it is checked textually against include/gingr_cuda.h: every exported function has a downcall handle,
and each handle's FunctionDescriptor has the header's arity and argument classes (int32 -> JAVA_INT, 64-bit integers ->
JAVA_LONG, double -> JAVA_DOUBLE, any pointer or array parameter -> ADDRESS).  The struct layouts are checked against
the ctypes mirrors (sizes and field order).  No GPU."""
import pathlib
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gingr_cuda.h")
SHIM = os.path.join(ROOT, "scala", "gingr", "api", "registration", "cuda", "GingrCudaNative.scala")


def _header_prototypes():
    h = pathlib.Path(HEADER).read_text()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    h = re.sub(r"//[^\n]*", "", h)
    out = {}
    for m in re.finditer(r"GINGR_API\s+([\w \*]+?)\s*\b(gingr_\w+)\s*\(([^;{]*?)\)\s*;", h):
        ret, name, args = m.group(1).strip(), m.group(2), " ".join(m.group(3).split())
        params = [] if args == "void" else [a.strip() for a in args.split(",")]
        out[name] = (_klass(ret + " x"), [_klass(p) for p in params])
    return out


def _klass(decl: str) -> str:
    if "*" in decl or "[" in decl:
        return "ADDRESS"
    t = decl.rsplit(None, 1)[0].replace("const", "").strip()
    return {"int32_t": "JAVA_INT", "uint32_t": "JAVA_INT", "int64_t": "JAVA_LONG", "uint64_t": "JAVA_LONG", "size_t": "JAVA_LONG",
            "double": "JAVA_DOUBLE"}[t]


def _shim_handles():
    s = pathlib.Path(SHIM).read_text()
    out = {}
    for m in re.finditer(r'fn\(\s*"(gingr_\w+)"\s*,([^)]*)\)', s):
        toks = [t.strip() for t in m.group(2).replace("\n", " ").split(",") if t.strip()]
        out[m.group(1)] = (toks[0], toks[1:])
    return s, out


def test_every_exported_function_has_a_matching_downcall_handle():
    protos = _header_prototypes()
    assert len(protos) >= 40
    _, handles = _shim_handles()
    assert sorted(set(protos) - set(handles)) == []
    assert sorted(set(handles) - set(protos)) == []
    for name, (ret, params) in protos.items():
        assert handles[name] == (ret, params), (name, handles[name], (ret, params))


def _layout_fields(src, name):
    body = src[src.index(f"val {name}: StructLayout"):]
    body = body[:body.index("\n  )")]
    fields = []
    for m in re.finditer(r'(?:sequenceLayout\((\d+),\s*(JAVA_\w+)\)|(JAVA_\w+))\.withName\("(\w+)"\)', body):
        n = int(m.group(1)) if m.group(1) else 1
        fields.append((m.group(4), m.group(2) or m.group(3), n))
    return fields


def test_struct_layouts_match_the_ctypes_mirrors():
    from gingr_b200 import _native as nat
    src, _ = _shim_handles()
    size = {"JAVA_INT": 4, "JAVA_DOUBLE": 8, "JAVA_LONG": 8}
    for layout, mirror in (("STATE", nat.GingrState), ("CONFIG", nat.GingrConfig), ("MCMC_SETTINGS", nat.GingrMcmcSettings)):
        fields = _layout_fields(src, layout)
        names = [f[0].rstrip("_") for f in mirror._fields_]
        assert [f[0] for f in fields] == names, (layout, fields, names)
        # FFM struct layouts carry no implicit padding: the listed members must already be naturally aligned and sum
        # to the C struct's size
        off = 0
        for fname, kind, n in fields:
            assert off % size[kind] == 0, (layout, fname)
            assert off == getattr(mirror, [f[0] for f in mirror._fields_][names.index(fname)]).offset, (layout, fname)
            off += size[kind] * n
        assert off == ctypes.sizeof(mirror), (layout, off, ctypes.sizeof(mirror))


def test_integration_notes_name_every_entry_point():
    doc = pathlib.Path(os.path.join(ROOT, "INTEGRATION.md")).read_text()
    assert [n for n in _header_prototypes() if n not in doc] == []


def _call_sites(src):
    """(handle name, number of arguments) of every `<handle>.invoke(...)` in a Scala source."""
    out = []
    for m in re.finditer(r"\b(\w+)\.invoke\(", src):
        i, depth, args, cur = m.end(), 1, 0, ""
        while depth:
            ch = src[i]
            if ch in "([{":
                depth += 1
            elif ch in ")]}":
                depth -= 1
                if depth == 0:
                    break
            elif ch == "," and depth == 1:
                args += 1
                cur = ""
                i += 1
                continue
            cur += ch
            i += 1
        out.append((m.group(1), args + (1 if cur.strip() else 0)))
    return out


def test_scala_call_sites_pass_the_declared_number_of_arguments():
    """CpdRegistrationCuda.scala calls the handles of GingrCudaNative.scala: every call site must pass as many arguments as
    the handle's FunctionDescriptor (and hence the C prototype) has parameters."""
    shim_src, handles = _shim_handles()
    by_val = {}
    for m in re.finditer(r'val\s+(\w+)\s*=\s*fn\(\s*"(gingr_\w+)"', shim_src):
        by_val[m.group(1)] = m.group(2)
    user = pathlib.Path(os.path.join(os.path.dirname(SHIM), "CpdRegistrationCuda.scala")).read_text()
    sites = _call_sites(user) + [s for s in _call_sites(shim_src) if s[0] in by_val]
    assert len(sites) >= 15
    for name, nargs in sites:
        assert name in by_val, f"{name}.invoke: no such handle"
        assert nargs == len(handles[by_val[name]][1]), (name, nargs, handles[by_val[name]][1])
