"""The Julia-side binding (volumefluxes.jl_b200/julia/B200Backend.jl) cannot be executed here (no julia binary), so it
is checked statically against the C header and the ctypes mirror: the SweB200Config struct layout, every ccall'ed
symbol and its arity, and the kernel-surface method signatures of backends/kernel_abstractions.jl:73, 83, 129."""
import ctypes as C
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
JL = (ROOT / "volumefluxes.jl_b200" / "julia" / "B200Backend.jl").read_text()
HEADER = (ROOT / "include" / "swe_b200.h").read_text()

JL_TYPES = {"Int32": (C.c_int32, 1), "Float64": (C.c_double, 1)}


def _julia_struct_fields(name):
    body = re.search(r"struct " + name + r"\n(.*?)\nend", JL, re.S).group(1)
    fields = []
    for decl in re.split(r"[;\n]", body):
        decl = decl.strip()
        if not decl or decl.startswith("#"):
            continue
        fname, ftype = [s.strip() for s in decl.split("::")]
        m = re.fullmatch(r"NTuple\{(\d+),\s*(\w+)\}", ftype)
        if m:
            fields.append((fname, JL_TYPES[m.group(2)][0], int(m.group(1))))
        else:
            fields.append((fname, JL_TYPES[ftype][0], 1))
    return fields


def _header_struct_fields(name):
    body = re.search(r"typedef struct " + name + r" \{(.*?)\} " + name + ";", HEADER, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for stmt in body.split(";"):
        stmt = " ".join(stmt.split())
        if not stmt:
            continue
        ctype, rest = stmt.split(" ", 1)
        ct = {"int32_t": C.c_int32, "double": C.c_double}[ctype]
        for item in rest.split(","):
            item = item.strip()
            m = re.fullmatch(r"(\w+)\[(\d+)\]", item)
            fields.append((m.group(1), ct, int(m.group(2))) if m else (item, ct, 1))
    return fields


def test_julia_config_struct_matches_header_and_ctypes(pkg):
    jl = _julia_struct_fields("SweB200Config")
    hd = _header_struct_fields("swe_b200_config")
    assert jl == hd, "SweB200Config in B200Backend.jl and swe_b200_config in include/swe_b200.h differ"
    py = []
    for fname, ftype in pkg.api._lib.Config._fields_:
        if hasattr(ftype, "_length_"):
            py.append((fname, ftype._type_, ftype._length_))
        else:
            py.append((fname, ftype, 1))
    assert py == hd, "the ctypes Config mirror and the header differ"


def _header_prototypes():
    protos = {}
    text = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    for m in re.finditer(r"\b(?:int|void\*|const char\*)\s+(swe_b200_\w+)\s*\(([^)]*)\)\s*;", text):
        args = [a.strip() for a in m.group(2).split(",") if a.strip() and a.strip() != "void"]
        protos[m.group(1)] = len(args)
    return protos


def test_every_julia_ccall_binds_a_declared_symbol_with_the_right_arity():
    protos = _header_prototypes()
    calls = re.findall(r"ccall\(\(:(swe_b200_\w+), libswe\),\s*[\w{}.]+,\s*\(([^)]*)\)", JL)
    assert len(calls) >= 20
    for sym, argtypes in calls:
        assert sym in protos, f"{sym} is ccall'ed in B200Backend.jl but not declared in include/swe_b200.h"
        depth, n, cur = 0, 0, ""
        for ch in argtypes + ",":          # count top-level commas (Ptr{Cvoid}, Ref{Int64}, ...)
            if ch == "{":
                depth += 1
            elif ch == "}":
                depth -= 1
            if ch == "," and depth == 0:
                n += 1 if cur.strip() else 0
                cur = ""
            else:
                cur += ch
        assert n == protos[sym], f"{sym}: {n} ccall argument types, the header declares {protos[sym]}"
    used = {s for s, _ in calls}
    for needed in ("swe_b200_create", "swe_b200_destroy", "swe_b200_set_state_interior", "swe_b200_get_state_interior",
                   "swe_b200_step", "swe_b200_simulate_to_time", "swe_b200_enumerate_grid", "swe_b200_kernel_family",
                   "swe_b200_update_bc", "swe_b200_reconstruct", "swe_b200_add_time_derivative",
                   "swe_b200_implicit_substep", "swe_b200_post_proc_substep"):
        assert needed in used, f"B200Backend.jl does not bind {needed}"


def test_julia_kernel_surface_has_the_reference_signatures():
    """backends/kernel_abstractions.jl:73, 83, 129"""
    assert re.search(r"function for_each_cell\(f, backend::B200Backend, grid, y\.\.\.\)", JL)
    assert re.search(r"function for_each_inner_cell\(f, backend::B200Backend, grid, direction, y\.\.\.; "
                     r"ghostcells=grid\.ghostcells\[direction\]\)", JL)
    assert re.search(r"function for_each_ghost_cell\(f, backend::B200Backend, grid, direction, y\.\.\.\)", JL)
    # the simulator is a genuine Simulator{...}: reference methods typed ::Simulator keep working
    assert "Simulator{typeof(backend),typeof(system),typeof(timestepper),typeof(grid),B200State{T},T}(" in JL
    for method in ("set_current_state!", "current_interior_state", "perform_step!", "simulate_to_time"):
        assert re.search(method.replace("!", r"\!") + r"\(sim::Simulator\{", JL), method
    # ConservedSystem allocates nothing on the host for this backend
    assert "create_buffer(backend::B200Backend{T}" in JL and "create_scalar(backend::B200Backend{T}" in JL
