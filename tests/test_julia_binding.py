"""Static check of the Julia binding (julia/PGBPB200.jl) against the C header (include/pgbp_b200.h).

Julia is not in the build image, so the wrapper cannot be executed here; this test parses it instead: the two C
structs must list the header's fields in the header's order with matching types, every `ccall` must name a declared
symbol with the declared number of arguments and ABI-compatible argument / return types, and the flag constants must
equal the header's #defines.  (The same ABI is exercised at run time by the ctypes mirror, tests/test_abi.py.)"""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, "include", "pgbp_b200.h")).read()
JULIA = open(os.path.join(ROOT, "julia", "PGBPB200.jl")).read()

C2J = {
    "int32_t": {"Int32"}, "int64_t": {"Int64"}, "uint32_t": {"UInt32"}, "size_t": {"Csize_t"}, "void": {"Cvoid"},
    "double*": {"Ptr{Float64}", "Ref{Float64}"}, "int32_t*": {"Ptr{Int32}", "Ref{Int32}"},
    "int64_t*": {"Ptr{Int64}", "Ref{Int64}"}, "uint8_t*": {"Ptr{UInt8}"}, "char*": {"Ptr{UInt8}", "Cstring"},
    "void*": {"Ptr{Cvoid}"}, "pgbp_plan*": {"Ptr{Cvoid}"}, "pgbp_batch*": {"Ptr{Cvoid}"}, "pgbp_comm*": {"Ptr{Cvoid}"},
    "pgbp_plan**": {"Ref{Ptr{Cvoid}}"}, "pgbp_batch**": {"Ref{Ptr{Cvoid}}"}, "pgbp_comm**": {"Ref{Ptr{Cvoid}}"},
    "double**": {"Ref{Ptr{Float64}}"}, "pgbp_plan_desc*": {"Ref{PlanDescC}", "Ptr{PlanDescC}"},
    "pgbp_family_table*": {"Ptr{FamilyTableC}"},
}


def strip_comments(txt):
    return re.sub(r"/\*.*?\*/", "", txt, flags=re.S)


def ctype(decl):
    """'const int32_t* name' -> 'int32_t*'"""
    d = re.sub(r"\bconst\b", "", decl).strip()
    d = re.sub(r"\s*/\*.*", "", d)
    m = re.match(r"^([A-Za-z_0-9]+)\s*(\**)\s*(?:const\s*)?(\**)\s*[A-Za-z_0-9]*(\[\d*\])?$", d)
    assert m, decl
    stars = m.group(2) + m.group(3) + ("*" if m.group(4) else "")
    return m.group(1) + stars


def header_functions():
    txt = strip_comments(HEADER)
    out = {}
    for m in re.finditer(r"\b(int32_t|int64_t)\s+(pgbp_\w+)\s*\(([^)]*)\)\s*;", txt):
        args = [a.strip() for a in m.group(3).replace("\n", " ").split(",")]
        args = [] if args == ["void"] else args
        out[m.group(2)] = (m.group(1), [ctype(a) for a in args])
    return out


def header_struct(name):
    txt = strip_comments(HEADER)
    m = re.search(r"typedef struct " + name + r" \{(.*?)\} " + name + ";", txt, flags=re.S)
    fields = []
    for line in m.group(1).split(";"):
        line = line.strip()
        if line:
            fields.append((re.findall(r"([A-Za-z_0-9]+)\s*$", line)[0], ctype(line)))
    return fields


def julia_struct(name):
    m = re.search(r"^struct " + name + r"\n(.*?)^end", JULIA, flags=re.S | re.M)
    fields = []
    for line in m.group(1).splitlines():
        line = line.split("#")[0].strip()
        if line:
            n, t = line.split("::")
            fields.append((n.strip(), t.strip()))
    return fields


def split_top(s):
    """split a Julia tuple body on top-level commas"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def julia_ccalls():
    calls = []
    for m in re.finditer(r"ccall\(\(:(\w+), LIB\),\s*(\w+),\s*\(", JULIA):
        start = m.end()
        depth, i = 1, start
        while depth:
            depth += {"(": 1, ")": -1}.get(JULIA[i], 0)
            i += 1
        calls.append((m.group(1), m.group(2), split_top(JULIA[start:i - 1])))
    return calls


def test_structs_match_the_header():
    for cname, jname in (("pgbp_family_table", "FamilyTableC"), ("pgbp_plan_desc", "PlanDescC")):
        hf, jf = header_struct(cname), julia_struct(jname)
        assert [n for n, _ in hf] == [n for n, _ in jf], (cname, hf, jf)
        for (n, ct), (_, jt) in zip(hf, jf):
            assert jt in C2J[ct], (cname, n, ct, jt)


def test_every_ccall_matches_its_prototype():
    protos = header_functions()
    calls = julia_ccalls()
    assert len(calls) >= 35
    for name, ret, args in calls:
        assert name in protos, name
        cret, cargs = protos[name]
        assert ret in C2J[cret], (name, ret, cret)
        assert len(args) == len(cargs), (name, args, cargs)
        for k, (jt, ct) in enumerate(zip(args, cargs)):
            assert jt in C2J[ct], (name, k, jt, ct)
    # the hot-path entry points of the header are all bound
    bound = {c[0] for c in calls}
    for must in ("pgbp_plan_create", "pgbp_batch_create_shared", "pgbp_assign_factors", "pgbp_assign_factors_ou",
                 "pgbp_assign_factors_device", "pgbp_calibrate", "pgbp_calibrate_async", "pgbp_propagate", "pgbp_integrate",
                 "pgbp_integrate_device", "pgbp_integrate_cov", "pgbp_factored_energy", "pgbp_factored_energy_device",
                 "pgbp_regularize_bycluster", "pgbp_regularize_onschedule", "pgbp_regularize_bynodesubtree",
                 "pgbp_reset_from_factors", "pgbp_factors_from_beliefs", "pgbp_reset_calibration_flags", "pgbp_set_belief",
                 "pgbp_get_belief", "pgbp_get_status", "pgbp_batch_set_stream", "pgbp_comm_create", "pgbp_comm_connect",
                 "pgbp_integrate_gather", "pgbp_comm_wait"):
        assert must in bound, must


def test_flag_constants_match_the_header():
    defs = dict(re.findall(r"#define\s+(PGBP_\w+)\s+(\d+)u?\b", HEADER))
    pairs = {"BATCH_FACTORS": "PGBP_BATCH_FACTORS", "BATCH_RESIDUALS": "PGBP_BATCH_RESIDUALS", "CAL_POSTORDER": "PGBP_CAL_POSTORDER",
             "CAL_PREORDER": "PGBP_CAL_PREORDER", "CAL_BOTH": "PGBP_CAL_BOTH", "CAL_RESIDNORM": "PGBP_CAL_RESIDNORM",
             "CAL_RESIDKLDIV": "PGBP_CAL_RESIDKLDIV", "CAL_AUTO": "PGBP_CAL_AUTO", "CAL_REFORDER": "PGBP_CAL_REFORDER",
             "PAIR_ZIP": "PGBP_PAIR_ZIP", "PAIR_PRODUCT": "PGBP_PAIR_PRODUCT"}
    for j, c in pairs.items():
        m = re.search(r"const " + j + r" = U?Int32\((\d+)\)", JULIA)
        assert m and m.group(1) == defs[c], (j, c)


def test_reference_signature_methods_exist():
    # the methods a call site written for ClusterGraphBelief needs (boundary hardening)
    for pat in (r"function PGBP\.assignfactors!\(b::BatchedClusterGraphBelief,\s*model::Union\{PGBP\.EvolutionaryModel",
                r"PGBP\.integratebelief!\(b::BatchedClusterGraphBelief, cgraph::MetaGraph, prenodes\)",
                r"function PGBP\.regularizebeliefs_bynodesubtree!\(b::BatchedClusterGraphBelief, cgraph::MetaGraph\)",
                r"function PGBP\.calibrate_optimize_cliquetree!\(b::BatchedClusterGraphBelief",
                r"function PGBP\.calibrate_optimize_clustergraph!\(b::BatchedClusterGraphBelief",
                r"function PGBP\.calibrate_exact_cliquetree!\(bi::BatchedClusterGraphBelief"):
        assert re.search(pat, JULIA), pat
