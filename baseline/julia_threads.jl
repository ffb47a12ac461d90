# baseline/julia_threads.jl -- the reference's own CPU path for BASELINE.json's metric: one ClusterGraphBelief per
# thread, `Threads.@threads` over trait replicates (north_star: "the reference's Julia CPU path ... timed on the GPU
# box's own host cores, core count stated").
#
#   julia --threads=auto --project=<env with PhyloGaussianBeliefProp> baseline/julia_threads.jl \
#         [--network lazaridis_2014.phy] [--ntraits 3] [--replicates 65536] [--seconds 10] [--steps 3]
#
# Julia is NOT part of the image this repository is built in (probed: `julia` absent), so this script has never run
# here; `bench.py --impl reference` calls it when a `julia` executable and an installed PhyloGaussianBeliefProp are
# found at run time and falls back to the C/OpenMP restatement (oracle/c, kind "port") otherwise.  It prints ONE
# JSON line {"value": calibrations/s, "cores": nthreads, "kind": "reference", "sample": ...}.
#
# Workload = BASELINE configs[1] as bench.py states it: lazaridis_2014 admixture graph, clique tree, MvFullBrownianMotion
# (p = 3, R = A A'/3 + 0.1 I, mu = 0, fixed root), per replicate: assignfactors! + calibrate!(beliefs, [spt]) +
# integratebelief! at the root cluster.
using PhyloGaussianBeliefProp, PhyloNetworks, Tables, Random, LinearAlgebra
const PGBP = PhyloGaussianBeliefProp

function arg(name, default)
    i = findfirst(==(name), ARGS)
    i === nothing ? default : ARGS[i+1]
end
netfile = arg("--network", joinpath(pkgdir(PGBP), "test", "example_networks", "lazaridis_2014.phy"))
p = parse(Int, arg("--ntraits", "3"))
B = parse(Int, arg("--replicates", "65536"))
seconds = parse(Float64, arg("--seconds", "10"))
steps = parse(Int, arg("--steps", "3"))

net = readnewick(netfile)
preorder!(net)
taxa = tiplabels(net)
rng = MersenneTwister(0xB200 + 2)
A = randn(rng, p, p)
R = Symmetric(A * A' / 3 + 0.1I)
model = PGBP.MvFullBrownianMotion(Matrix(R), zeros(p))          # fixed root
ct = PGBP.clustergraph!(net, PGBP.Cliquetree())
spt = PGBP.spanningtree_clusterlist(ct, net.vec_node)
rootj = spt[3][1]

"trait replicate simulated down the network: X_v = sum_k gamma_k X_pa_k + N(0, sum_k gamma_k^2 t_k R)"
function simulate(rng)
    L = cholesky(R).L
    X = Dict{Int,Vector{Float64}}()
    for node in net.vec_node
        if node === net.vec_node[1]
            X[node.number] = zeros(p); continue
        end
        m = zeros(p); v = 0.0
        for e in node.edge
            getchild(e) === node || continue
            m .+= e.gamma .* X[getparent(e).number]; v += e.gamma^2 * e.length
        end
        X[node.number] = m .+ sqrt(v) .* (L * randn(rng, p))
    end
    tips = [X[n.number] for n in net.leaf]
    NamedTuple{Tuple(Symbol("x$t") for t in 1:p)}(Tuple([tip[t] for tip in tips] for t in 1:p))
end

nthr = Threads.nthreads()
# one set of beliefs per thread (allocated once, as a user of the package would)
tbl0 = simulate(rng)
work = map(1:nthr) do _
    b, (n2c, n2fam, n2fix, n2d, c2n) = PGBP.allocatebeliefs(tbl0, taxa, net.vec_node, ct, model)
    PGBP.ClusterGraphBelief(b, n2c, n2fam, n2fix, c2n)
end

function run!(tbls)
    ll = Vector{Float64}(undef, length(tbls))
    Threads.@threads :static for i in eachindex(tbls)
        cgb = work[Threads.threadid()]
        PGBP.assignfactors!(cgb.belief, model, tbls[i], taxa, net.vec_node, cgb.node2cluster, cgb.node2family, cgb.node2fixed)
        PGBP.init_messagecalibrationflags_reset!(cgb, false)
        PGBP.calibrate!(cgb, [spt])
        ll[i] = PGBP.integratebelief!(cgb, rootj)[2]
    end
    ll
end

# bounded sample of the B replicates: sized from a probe so that one step lasts ~ `seconds`
probe = [simulate(rng) for _ in 1:min(B, 64 * nthr)]
run!(probe)                                   # compile + warm up
t = @elapsed run!(probe)
n = clamp(round(Int, length(probe) / t * seconds), length(probe), B)
tbls = n == length(probe) ? probe : vcat(probe, [simulate(rng) for _ in 1:(n - length(probe))])
run!(tbls)
dt = minimum(@elapsed(run!(tbls)) for _ in 1:steps)
println("{\"value\": $(n / dt), \"unit\": \"calibrations/s\", \"cores\": $nthr, \"kind\": \"reference\", ",
        "\"sample\": \"$n of $B replicates per step, best of $steps (assignfactors! + calibrate! + integratebelief!, Threads.@threads)\", ",
        "\"seconds_per_step\": $dt}")
