# PGBPB200.jl -- reference-side binding of libpgbp_b200.so (include/pgbp_b200.h).
#
# This is the host code a maintainer of PhyloGaussianBeliefProp.jl adds to use the B200 library
# as a drop-in for the message-passing path: a batched belief container plus methods of the
# package's own function names (`calibrate!`, `propagate_1traversal_postorder!`,
# `propagate_belief!`, `integratebelief!`, `factored_energy`, `assignfactors!` /
# `init_factors_frommodel!`, `regularizebeliefs_*!`, `init_beliefs_reset_fromfactors!`, ...), each
# a thin `ccall`.  Everything graph-related (cluster graphs, spanning-tree schedules, models,
# optimisers) stays what the package already does.
#
# STATUS: Julia is not available in the image this repository is built and tested in, so this file
# has never been executed here.  The same boundary is exercised by the Python ctypes mirror
# (phylogaussianbeliefprop.jl_b200/api.py), which the parity tests drive; this module is a
# line-for-line counterpart of that mirror.  `PGBPB200.dumpplan(...)` writes the plan arrays as
# JSON so they can be diffed against `workloads/*.json` wherever Julia exists.
module PGBPB200

import PhyloGaussianBeliefProp as PGBP
using PhyloGaussianBeliefProp: ClusterGraphBelief, CanonicalBelief, nclusters, nsepsets, scopeindex,
    sepsetindex, clusterindex
import PhyloNetworks as PN
using MetaGraphsNext: labels, edge_labels

const LIB = get(ENV, "PGBP_B200_LIB", joinpath(@__DIR__, "..", "phylogaussianbeliefprop.jl_b200", "lib", "libpgbp_b200.so"))

# ---------------------------------------------------------------- error handling
struct PgbpError <: Exception
    code::Int32
    msg::String
end
function check(rc::Int32)
    rc == 0 && return nothing
    buf = Vector{UInt8}(undef, 512)
    ccall((:pgbp_last_error, LIB), Int32, (Ptr{UInt8}, Csize_t), buf, 512)
    throw(PgbpError(rc, unsafe_string(pointer(buf))))
end

# flags (mirror include/pgbp_b200.h)
const BATCH_FACTORS = UInt32(1); const BATCH_RESIDUALS = UInt32(2)
const CAL_POSTORDER = UInt32(1); const CAL_PREORDER = UInt32(2); const CAL_BOTH = UInt32(3)
const CAL_RESIDNORM = UInt32(4); const CAL_RESIDKLDIV = UInt32(8); const CAL_AUTO = UInt32(16)
const PAIR_ZIP = Int32(0); const PAIR_PRODUCT = Int32(1)

# ---------------------------------------------------------------- C structs
struct FamilyTableC
    nnodes::Int32
    ntips::Int32
    node_cluster::Ptr{Int32}
    mem_off::Ptr{Int32}
    mem_pos::Ptr{Int32}
    mem_length::Ptr{Float64}
    mem_gamma::Ptr{Float64}
    mem_color::Ptr{Int32}
    node_datarow::Ptr{Int32}
    root_fixed::Int32
    mem_tpos::Ptr{Int32}     # trait-level scopes (missing data); C_NULL = full scopes
    tip_missing::Ptr{UInt8}
end
struct PlanDescC
    nclusters::Int32
    nsepsets::Int32
    ntraits::Int32
    belief_dim::Ptr{Int32}
    sepset_clusters::Ptr{Int32}
    upind_off::Ptr{Int32}
    upind::Ptr{Int32}
    ntrees::Int32
    tree_off::Ptr{Int32}
    tree_parent::Ptr{Int32}
    tree_child::Ptr{Int32}
    families::Ptr{FamilyTableC}
end

# ---------------------------------------------------------------- plan (static index work, once per graph)
"""
    PlanArrays(beliefs, nclusters, cgraph, schedule; families=nothing)

Index data of one cluster graph, exactly as the package produces it:
`beliefs` from `allocatebeliefs` (clusters then sepsets), `schedule` from
`spanningtree(s)_clusterlist`.  All indices converted to 0-based Int32.
"""
struct PlanArrays
    nclusters::Int32
    ntraits::Int32
    belief_dim::Vector{Int32}
    sepset_clusters::Vector{Int32}
    upind_off::Vector{Int32}
    upind::Vector{Int32}
    tree_off::Vector{Int32}
    tree_parent::Vector{Int32}
    tree_child::Vector{Int32}
    families::Union{Nothing,NamedTuple}
end

function PlanArrays(beliefs::AbstractVector, nclu::Integer, cgraph, schedule::AbstractVector; families=nothing)
    lab2idx = Dict(l => Int32(i - 1) for (i, l) in enumerate(labels(cgraph)))
    dims = Int32[PGBP.dimension(b) for b in beliefs]  # sum(inscope), src/beliefs.jl:291
    sc = Int32[]; off = Int32[0]; up = Int32[]
    for s in beliefs[nclu+1:end]
        (l1, l2) = s.metadata
        a, b = lab2idx[l1], lab2idx[l2]
        push!(sc, a, b)
        append!(up, Int32.(scopeindex(s, beliefs[a+1]) .- 1)); push!(off, length(up))
        append!(up, Int32.(scopeindex(s, beliefs[b+1]) .- 1)); push!(off, length(up))
    end
    toff = Int32[0]; tp = Int32[]; tc = Int32[]
    for spt in schedule            # (parent_labels, child_labels, parent_indices, child_indices)
        append!(tp, Int32.(spt[3] .- 1)); append!(tc, Int32.(spt[4] .- 1)); push!(toff, length(tp))
    end
    PlanArrays(nclu, beliefs[1].ntraits, dims, sc, off, up, toff, tp, tc, families)
end

"""
    familiestable(prenodes, node2cluster, node2family, node2fixed, beliefs, taxa; edgecolor = e -> 0, rootfixed, tbl = nothing)

Node-family table for device-side `assignfactors!` (Brownian-motion models).  With `tbl` (the column
table the beliefs were allocated from) the missingness pattern of the tips is recorded; whenever a
value is missing or a belief has a partial trait scope the table carries trait-level scope positions
and the library runs the reference's absorb / marginalise sequence per node family on the device.
"""
function familiestable(prenodes, node2cluster, node2family, node2fixed, beliefs, taxa; edgecolor = e -> 0, rootfixed::Bool, tbl = nothing)
    p = beliefs[1].ntraits
    mem_off = Int32[0]; mem_pos = Int32[]; mem_len = Float64[]; mem_gam = Float64[]; mem_col = Int32[]; row = Int32[]
    mem_tpos = Int32[]; partial = false
    for (v, node) in enumerate(prenodes)
        be = beliefs[node2cluster[v]]
        nd = vec(sum(be.inscope, dims=1)); cs = cumsum(vcat(0, nd))
        for (k, q) in enumerate(node2family[v])
            if node2fixed[q]
                push!(mem_pos, -1); append!(mem_tpos, fill(Int32(-1), p))
            else
                jj = findfirst(isequal(q), be.nodelabel)
                nd[jj] == p || (partial = true)
                push!(mem_pos, cs[jj])
                run = cs[jj]
                for t in 1:p
                    if be.inscope[t, jj]; push!(mem_tpos, run); run += 1 else push!(mem_tpos, -1) end
                end
            end
            if k == 1
                push!(mem_len, 0.0); push!(mem_gam, 1.0); push!(mem_col, 0)
            else
                e = first(e for e in node.edge if PN.getchild(e) === node && PN.getparent(e) === prenodes[q])
                push!(mem_len, e.length); push!(mem_gam, e.gamma); push!(mem_col, edgecolor(e))
            end
        end
        push!(mem_off, length(mem_pos))
        push!(row, node.leaf ? findfirst(isequal(node.name), taxa) - 1 : -1)
    end
    # tip_missing[row, trait] (row-major, like a tip-data record)
    miss = tbl === nothing ? zeros(UInt8, p * length(taxa)) :
           UInt8[ismissing(col[i]) for i in eachindex(taxa) for col in tbl]
    scoped = partial || any(!iszero, miss)
    (nnodes = Int32(length(prenodes)), ntips = Int32(length(taxa)), root_fixed = Int32(rootfixed),
     node_cluster = Int32.(node2cluster .- 1), mem_off = mem_off, mem_pos = mem_pos, mem_length = mem_len,
     mem_gamma = mem_gam, mem_color = mem_col, node_datarow = row,
     mem_tpos = scoped ? mem_tpos : Int32[], tip_missing = scoped ? miss : UInt8[])
end

mutable struct Plan
    handle::Ptr{Cvoid}
    arrays::PlanArrays
    function Plan(a::PlanArrays)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        f = a.families
        GC.@preserve a begin
            famref = Ref{FamilyTableC}()
            famptr = Ptr{FamilyTableC}(C_NULL)
            if f !== nothing
                famref[] = FamilyTableC(f.nnodes, f.ntips, pointer(f.node_cluster), pointer(f.mem_off), pointer(f.mem_pos),
                                        pointer(f.mem_length), pointer(f.mem_gamma), pointer(f.mem_color),
                                        pointer(f.node_datarow), f.root_fixed,
                                        isempty(f.mem_tpos) ? Ptr{Int32}(C_NULL) : pointer(f.mem_tpos),
                                        isempty(f.tip_missing) ? Ptr{UInt8}(C_NULL) : pointer(f.tip_missing))
                famptr = Base.unsafe_convert(Ptr{FamilyTableC}, famref)
            end
            d = Ref(PlanDescC(a.nclusters, length(a.belief_dim) - a.nclusters, a.ntraits, pointer(a.belief_dim),
                              pointer(a.sepset_clusters), pointer(a.upind_off), pointer(a.upind),
                              length(a.tree_off) - 1, pointer(a.tree_off), pointer(a.tree_parent), pointer(a.tree_child), famptr))
            GC.@preserve famref d check(ccall((:pgbp_plan_create, LIB), Int32, (Ref{PlanDescC}, Ref{Ptr{Cvoid}}), d, h))
        end
        p = new(h[], a)
        finalizer(x -> ccall((:pgbp_plan_destroy, LIB), Int32, (Ptr{Cvoid},), x.handle), p)
    end
end

# ---------------------------------------------------------------- batched beliefs
"""
    BatchedClusterGraphBelief(plan, B; device=0, factors=true, residuals=true)

`B` independent replicas (trait replicates and/or parameter vectors) of a
`ClusterGraphBelief` (src/clustergraphbeliefs.jl:26-53), resident on one GPU.
"""
mutable struct BatchedClusterGraphBelief
    handle::Ptr{Cvoid}
    plan::Plan
    B::Int
    schedule::Vector            # the spanning trees the plan was built with (for tree ids)
    # sharedgroup = g > 1: the g consecutive elements of a group are trait replicates under one parameter
    # vector; their (identical) precisions J are stored and updated once per group
    function BatchedClusterGraphBelief(plan::Plan, B::Integer, schedule; device::Integer=0, factors=true, residuals=true,
                                       sharedgroup::Integer=0)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        fl = (factors ? BATCH_FACTORS : UInt32(0)) | (residuals ? BATCH_RESIDUALS : UInt32(0))
        check(ccall((:pgbp_batch_create_shared, LIB), Int32, (Ptr{Cvoid}, Int64, Int64, Int32, UInt32, Ref{Ptr{Cvoid}}),
                    plan.handle, B, sharedgroup, device, fl, h))
        b = new(h[], plan, B, collect(schedule))
        finalizer(x -> ccall((:pgbp_batch_destroy, LIB), Int32, (Ptr{Cvoid},), x.handle), b)
    end
end
PGBP.nclusters(b::BatchedClusterGraphBelief) = Int(b.plan.arrays.nclusters)
PGBP.nsepsets(b::BatchedClusterGraphBelief) = length(b.plan.arrays.belief_dim) - PGBP.nclusters(b)
dimension(b::BatchedClusterGraphBelief, j::Integer) = Int(b.plan.arrays.belief_dim[j])

treeid(b::BatchedClusterGraphBelief, spt) = Int32(findfirst(t -> t[3] == spt[3] && t[4] == spt[4], b.schedule) - 1)

# belief access: J (m,m,B), h (m,B), g (B,) -- Julia's column-major arrays map directly
function setbelief!(b::BatchedClusterGraphBelief, j::Integer, J::Array{Float64,3}, h::Matrix{Float64}, g::Vector{Float64})
    GC.@preserve J h g check(ccall((:pgbp_set_belief, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                                   b.handle, j - 1, J, h, g))
end
function getbelief(b::BatchedClusterGraphBelief, j::Integer)
    m = dimension(b, j)
    J = Array{Float64}(undef, m, m, b.B); h = Matrix{Float64}(undef, m, b.B); g = Vector{Float64}(undef, b.B)
    GC.@preserve J h g check(ccall((:pgbp_get_belief, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                                   b.handle, j - 1, J, h, g))
    return J, h, g
end
function status(b::BatchedClusterGraphBelief)
    st = Vector{Int32}(undef, b.B)
    GC.@preserve st check(ccall((:pgbp_get_status, LIB), Int32, (Ptr{Cvoid}, Ptr{Int32}), b.handle, st))
    st
end

# ---------------------------------------------------------------- the package's API, batched
"""
    assignfactors!(b, params, tipdata; ncolors=1, pairing=:zip)

Device-side `assignfactors!` (src/beliefs.jl:786-861) for Brownian-motion models.
`params`: (ncolors*p*p + p + p*p, nparamsets) -- rates `R_c`, root mean, root variance per column;
`tipdata`: (p, ntips, ndatasets).
"""
function PGBP.assignfactors!(b::BatchedClusterGraphBelief, params::Matrix{Float64}, tipdata::Array{Float64,3};
                             ncolors::Integer=1, pairing::Symbol=:zip)
    GC.@preserve params tipdata check(ccall((:pgbp_assign_factors, LIB), Int32,
        (Ptr{Cvoid}, Int32, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Int32),
        b.handle, ncolors, params, size(params, 2), tipdata, size(tipdata, 3), pairing === :product ? PAIR_PRODUCT : PAIR_ZIP))
end
const init_factors_frommodel! = PGBP.assignfactors!   # spelling of the revision named in BASELINE.json

"pack one Brownian-motion parameter set (`R` per colour, root mean, root variance; 0 = fixed root)"
bmparams(rates::Vector{<:AbstractMatrix}, μ::AbstractVector, v::AbstractMatrix=zeros(length(μ), length(μ))) =
    vcat((vec(Matrix{Float64}(R)) for R in rates)..., Float64.(μ), vec(Matrix{Float64}(v)))
bmparams(m::PGBP.MvFullBrownianMotion) = bmparams([Matrix(m.R)], m.μ, Matrix(m.v))
bmparams(m::PGBP.MvDiagBrownianMotion) = bmparams([Matrix(PGBP.LA.Diagonal(m.R))], m.μ, Matrix(PGBP.LA.Diagonal(m.v)))
bmparams(m::PGBP.UnivariateBrownianMotion) = bmparams([fill(m.σ2, 1, 1)], [m.μ], fill(m.v, 1, 1))

function _flags(update_residualnorm, update_residualkldiv, auto)
    (update_residualnorm ? CAL_RESIDNORM : UInt32(0)) | (update_residualkldiv ? CAL_RESIDKLDIV : UInt32(0)) |
    (auto ? CAL_AUTO : UInt32(0))
end

"""
    calibrate!(b::BatchedClusterGraphBelief, schedule, niter=1; auto, info, update_residualnorm, update_residualkldiv)

Same semantics as `calibrate!` (src/calibration.jl:35-60), per batch element:
returns `(succ::BitVector, iscal::BitVector)`.
"""
function PGBP.calibrate!(b::BatchedClusterGraphBelief, schedule::AbstractVector, niter::Integer=1;
                         auto::Bool=false, info::Bool=false, verbose::Bool=true,
                         update_residualnorm::Bool=true, update_residualkldiv::Bool=false, direction::UInt32=CAL_BOTH)
    ids = Int32[treeid(b, spt) for spt in schedule]
    succ = Vector{Int32}(undef, b.B); iscal = Vector{Int32}(undef, b.B)
    it = info ? Matrix{Int32}(undef, 2, b.B) : nothing
    GC.@preserve ids succ iscal it check(ccall((:pgbp_calibrate, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Int32}, Int32, Int32, UInt32, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}),
        b.handle, ids, length(ids), niter, direction | _flags(update_residualnorm, update_residualkldiv, auto),
        succ, iscal, info ? pointer(it) : C_NULL))
    if info
        for e in 1:b.B
            it[1, e] > 0 && verbose && @info "element $e: calibration reached: iteration $(it[1,e]), schedule tree $(it[2,e])"
        end
    end
    return succ .!= 0, iscal .!= 0
end
PGBP.propagate_1traversal_postorder!(b::BatchedClusterGraphBelief, spt...; kw...) =
    PGBP.calibrate!(b, [spt]; direction=CAL_POSTORDER, kw...)[1]
PGBP.propagate_1traversal_preorder!(b::BatchedClusterGraphBelief, spt...; kw...) =
    PGBP.calibrate!(b, [spt]; direction=CAL_PREORDER, kw...)[1]

"`propagate_belief!(cluster_to, sepset, cluster_from, residual)` (src/beliefupdates.jl:634) by belief index"
PGBP.propagate_belief!(b::BatchedClusterGraphBelief, to::Integer, sepset::Integer, from::Integer) =
    check(ccall((:pgbp_propagate, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Int32, UInt32), b.handle, from - 1, sepset - 1, to - 1, 0))

"`integratebelief!(beliefs, j)` -> `(μ::Matrix (m,B), norm::Vector (B))` (src/clustergraphbeliefs.jl:194)"
function PGBP.integratebelief!(b::BatchedClusterGraphBelief, j::Integer)
    μ = Matrix{Float64}(undef, dimension(b, j), b.B); nrm = Vector{Float64}(undef, b.B)
    GC.@preserve μ nrm check(ccall((:pgbp_integrate, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}), b.handle, j - 1, μ, nrm))
    return μ, nrm
end

"`integratebelief!` + `inv(J)`: `(μ (m,B), cov (m,m,B), norm (B))`, the conditional moments of calibrate_exact_cliquetree!"
function integratebelief_cov!(b::BatchedClusterGraphBelief, j::Integer)
    m = dimension(b, j)
    μ = Matrix{Float64}(undef, m, b.B); cov = Array{Float64}(undef, m, m, b.B); nrm = Vector{Float64}(undef, b.B)
    GC.@preserve μ cov nrm check(ccall((:pgbp_integrate_cov, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                                       b.handle, j - 1, μ, cov, nrm))
    return μ, cov, nrm
end

"`factored_energy(beliefs)` -> (3,B) matrix: average energy, approximate entropy, factored energy (src/score.jl:151)"
function PGBP.factored_energy(b::BatchedClusterGraphBelief)
    out = Matrix{Float64}(undef, 3, b.B)
    GC.@preserve out check(ccall((:pgbp_factored_energy, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), b.handle, out))
    out
end
function PGBP.free_energy(b::BatchedClusterGraphBelief)
    out = PGBP.factored_energy(b); out[3, :] .*= -1; out
end

PGBP.regularizebeliefs_bycluster!(b::BatchedClusterGraphBelief, cgraph=nothing) =
    check(ccall((:pgbp_regularize_bycluster, LIB), Int32, (Ptr{Cvoid},), b.handle))
PGBP.regularizebeliefs_onschedule!(b::BatchedClusterGraphBelief, cgraph=nothing) =
    check(ccall((:pgbp_regularize_onschedule, LIB), Int32, (Ptr{Cvoid},), b.handle))
PGBP.init_beliefs_reset_fromfactors!(b::BatchedClusterGraphBelief) =
    check(ccall((:pgbp_reset_from_factors, LIB), Int32, (Ptr{Cvoid},), b.handle))
PGBP.init_factors_frombeliefs!(b::BatchedClusterGraphBelief) =
    check(ccall((:pgbp_factors_from_beliefs, LIB), Int32, (Ptr{Cvoid},), b.handle))
PGBP.init_messagecalibrationflags_reset!(b::BatchedClusterGraphBelief, reset_kl::Bool=true) =
    check(ccall((:pgbp_reset_calibration_flags, LIB), Int32, (Ptr{Cvoid}, Int32), b.handle, reset_kl))

"""
    regularizebeliefs_bynodesubtree!(b, cgraph)

Index program of the loop at src/clustergraphbeliefs.jl:314-340 (which clusters define a node's ϵ,
which (cluster, sepset) pairs and diagonal positions it is added to), then one library call.
"""
function PGBP.regularizebeliefs_bynodesubtree!(b::BatchedClusterGraphBelief, beliefs::ClusterGraphBelief, cgraph)
    eo = Int32[0]; ec = Int32[]; so = Int32[0]; sc = Int32[]; ss = Int32[]; io = Int32[0]; ic = Int32[]; is = Int32[]
    for (node_ind, (nodelab, clusterlabs)) in enumerate(PGBP.get_nodesymbols2index(cgraph) |> pairs)
        # follows the reference: clusters containing the node, then its node subtree in preorder
        sch = PGBP.nodesubtree_clusterlist(PGBP.nodesubtree(cgraph, nodelab, node_ind)[1], nodelab)
        isempty(sch[3]) && (push!(eo, length(ec)); push!(so, length(sc)); continue)
        cl = unique(vcat(sch[3], sch[4]))
        append!(ec, Int32.(cl .- 1)); push!(eo, length(ec))
        for (pl, chl) in zip(sch[1], sch[2])
            ci = clusterindex(chl, beliefs); si = sepsetindex(pl, chl, beliefs)
            (s_ind, c_ind) = scopeindex(node_ind, beliefs.belief[si], beliefs.belief[ci])
            push!(sc, ci - 1); push!(ss, si - 1 - PGBP.nclusters(beliefs))
            append!(ic, Int32.(c_ind .- 1)); append!(is, Int32.(s_ind .- 1)); push!(io, length(ic))
        end
        push!(so, length(sc))
    end
    GC.@preserve eo ec so sc ss io ic is check(ccall((:pgbp_regularize_bynodesubtree, LIB), Int32,
        (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}),
        b.handle, length(eo) - 1, eo, ec, so, sc, ss, io, ic, is))
end

# ---------------------------------------------------------------- plan dump, for diffing against workloads/*.json
function dumpplan(io::IO, a::PlanArrays)
    println(io, "{\"nclusters\": $(a.nclusters), \"ntraits\": $(a.ntraits), \"belief_dim\": $(a.belief_dim), ",
            "\"sepset_clusters\": $(a.sepset_clusters), \"upind_off\": $(a.upind_off), \"upind\": $(a.upind), ",
            "\"tree_off\": $(a.tree_off), \"tree_parent\": $(a.tree_parent), \"tree_child\": $(a.tree_child)}")
end

end # module
