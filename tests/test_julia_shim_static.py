"""The Julia shim (julia/MolecularDynamicsB200.jl) cannot be executed in this image (no Julia), so it is checked STATICALLY
against include/mdb200.h: every `ccall` must name an entry point the header declares, with the same number of arguments
and position by position a compatible type, and the mirrored structs must list the header's fields in the header's order
with matching widths.  This is the drift guard for the reference-side binding (VERDICT r01 "shim drift")."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mdb200.h")
SHIM = os.path.join(ROOT, "julia", "MolecularDynamicsB200.jl")


def strip_c_comments(text):
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    return re.sub(r"//[^\n]*", " ", text)


def c_kind(decl):
    """category of one C parameter / field declaration"""
    d = re.sub(r"\bconst\b", " ", decl).strip()
    array = re.search(r"\[(\d*)\]\s*$", d)
    if array:
        d = d[:array.start()].strip()
    stars = d.count("*")
    d = d.replace("*", " ")
    toks = d.split()
    # drop the parameter name (last identifier) when a type remains in front of it
    base = " ".join(toks[:-1]) if len(toks) > 1 else toks[0]
    base = {"unsigned long long": "uint64_t", "mdb_engine_s": "mdb_handle"}.get(base, base)
    if array:
        stars += 1
    if base == "mdb_handle":
        return "handle" if stars == 0 else "handle*"
    if base == "char":
        return "cstring" if stars == 1 else "char"
    if base == "void":
        return "void" + "*" * stars
    scalar = {"int": "i32", "int32_t": "i32", "int64_t": "i64", "uint64_t": "u64", "double": "f64", "float": "f32"}.get(base, "struct:" + base)
    return scalar + "*" * stars


def header_functions():
    text = strip_c_comments(open(HEADER).read())
    out = {}
    for m in re.finditer(r"\b(int|const\s+char\s*\*)\s*(mdb_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text):
        args = [a.strip() for a in m.group(3).split(",")] if m.group(3).strip() not in ("", "void") else []
        out[m.group(2)] = ("cstring" if "char" in m.group(1) else "i32", [c_kind(a) for a in args])
    return out


def header_struct(name):
    text = strip_c_comments(open(HEADER).read())
    m = re.search(r"typedef\s+struct\s+%s\s*\{(.*?)\}\s*%s\s*;" % (name, name), text, flags=re.S)
    assert m, name
    fields = []
    for stmt in m.group(1).split(";"):
        stmt = stmt.strip()
        if not stmt:
            continue
        typ, rest = stmt.split(None, 1)
        for item in rest.split(","):
            item = item.strip()
            arr = re.match(r"([A-Za-z0-9_]+)\s*\[(\d+)\]", item)
            if arr:
                fields.append((arr.group(1), c_kind(typ + " x"), int(arr.group(2))))
            else:
                fields.append((item, c_kind(typ + " x"), 1))
    return fields


JL_SCALAR = {"Int32": "i32", "Cint": "i32", "Int64": "i64", "UInt64": "u64", "Float64": "f64", "Cdouble": "f64"}


def jl_kind(t):
    t = t.strip()
    if t == "Handle":
        return "handle"
    if t == "Cstring":
        return "cstring"
    if t in JL_SCALAR:
        return JL_SCALAR[t]
    m = re.match(r"(Ptr|Ref)\{(.+)\}$", t)
    if m:
        inner = m.group(2).strip()
        if inner == "Handle":
            return "handle*"
        if inner == "Cvoid":
            return "void*"
        if inner in ("UInt8", "Cchar"):
            return "cstring"
        if inner in JL_SCALAR:
            return JL_SCALAR[inner] + "*"
        if re.match(r"Ptr\{", inner):
            return jl_kind(inner) + "*"
        return "struct:" + inner + "*"
    raise AssertionError("unmapped Julia type %r" % t)


def split_top(s):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "{(":
            depth += 1
        elif ch in "})":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        parts.append(cur)
    return [p.strip() for p in parts]


def shim_ccalls():
    text = re.sub(r"#[^\n]*", "", open(SHIM).read())
    calls = []
    for m in re.finditer(r"ccall\(\(:(mdb_[a-z0-9_]+),\s*libmdb\),\s*([A-Za-z0-9_]+),\s*\(", text):
        # argument type tuple: balanced parentheses from here
        i, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(text[i], 0)
            i += 1
        calls.append((m.group(1), m.group(2), split_top(text[m.end():i - 1])))
    return calls


def shim_struct(name):
    text = open(SHIM).read()
    m = re.search(r"^struct\s+%s\s*\n(.*?)^end" % name, text, flags=re.S | re.M)
    assert m, name
    fields = []
    for line in m.group(1).splitlines():
        line = re.sub(r"#.*", "", line).strip()
        if not line:
            continue
        fname, ftype = [p.strip() for p in line.split("::")]
        tup = re.match(r"NTuple\{(\d+),\s*([A-Za-z0-9]+)\}", ftype)
        if tup:
            fields.append((fname, JL_SCALAR[tup.group(2)], int(tup.group(1))))
        else:
            fields.append((fname, JL_SCALAR[ftype], 1))
    return fields


STRUCT_NAMES = {"struct:MdbConfig*": "struct:mdb_config*", "struct:FireParams*": "struct:mdb_fire_params*"}


def compatible(jl, c):
    jl = STRUCT_NAMES.get(jl, jl)
    if jl == c:
        return True
    # a C `double out[4]` / `char id[128]` parameter is a pointer; Julia passes arrays as Ptr / Ref of the element type
    if c == "char*" and jl in ("cstring", "void*"):
        return True
    if c == "cstring" and jl in ("cstring",):
        return True
    if c.startswith("void*") and jl.endswith("*"):
        return True
    return False


def test_every_ccall_matches_a_header_prototype():
    funcs = header_functions()
    assert len(funcs) >= 38, sorted(funcs)
    calls = shim_ccalls()
    assert len(calls) >= 20
    for name, ret, args in calls:
        assert name in funcs, "%s is not declared in include/mdb200.h" % name
        cret, cargs = funcs[name]
        assert jl_kind(ret) == cret or (cret == "cstring" and ret == "Cstring"), (name, ret, cret)
        assert len(args) == len(cargs), "%s: %d arguments in the shim, %d in the header" % (name, len(args), len(cargs))
        for k, (a, c) in enumerate(zip(args, cargs)):
            assert compatible(jl_kind(a), c), "%s argument %d: Julia %s vs C %s" % (name, k, a, c)


def test_mirrored_structs_follow_the_header():
    for jl_name, c_name in (("MdbConfig", "mdb_config"), ("FireParams", "mdb_fire_params")):
        jl, c = shim_struct(jl_name), header_struct(c_name)
        # the header may fold trailing padding into a reserved array; compare the flattened scalar sequences
        flat = lambda fields: [k for _, k, n in fields for _ in range(n)]
        assert flat(jl) == flat(c), (jl_name, jl, c)
        named_jl = [f for f, _, _ in jl if not f.startswith("reserved")]
        named_c = [f for f, _, _ in c if not f.startswith("reserved")]
        assert named_jl == named_c, (named_jl, named_c)


def test_the_drop_in_entry_points_are_all_bound():
    """what run_simulation! needs end to end must be reachable from the shim"""
    bound = {name for name, _, _ in shim_ccalls()}
    for need in ("mdb_create", "mdb_destroy", "mdb_upload", "mdb_download", "mdb_set_velocities", "mdb_run_nve", "mdb_run_nvt", "mdb_run_brownian",
                 "mdb_compute_forces", "mdb_thermo", "mdb_last_error", "mdb_init_velocities", "mdb_fire_minimize", "mdb_frame_capture",
                 "mdb_frame_write_lammps", "mdb_frame_flush", "mdb_checkpoint_save", "mdb_checkpoint_load", "mdb_upload_owned", "mdb_download_owned"):
        assert need in bound, need
