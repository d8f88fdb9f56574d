# SNaQB200.jl — the shim a SNaQ.jl maintainer would add to route the quartet pseudolikelihood hot path
# through libsnaqb200.so (include/snaq_b200.h).  NOT TESTED HERE: there is no Julia in the build image.
# It touches nothing but the call sites listed in INTEGRATION.md; the public API (snaq!, optBL!,
# topologyQpseudolik!, bootsnaq, readtableCF, DataCF, HybridNetwork) is unchanged.
#
#   using SNaQ, SNaQB200
#   SNaQB200.install!()                  # SNaQ.optBL!, topologyQpseudolik!, sampleEdgeQuartetWeighted, readtableCF! now use the engine
#   net = snaq!(startnet, d; hmax = 1)   # unchanged user code; one engine per (worker process, DataCF), GPU (myid()-1) % 8
module SNaQB200

using SNaQ, PhyloNetworks
import Distributed
import SNaQ: istIdentifiable, fromBadDiamondI, inCycle, hasHybEdge, isBadDiamondI, isBadDiamondII, isBadTriangle,
             gammaz, numBad, ht, numht, DataCF, Quartet

const LIB = get(ENV, "SNAQ_B200_LIB", "libsnaqb200.so")

struct FlatNet            # mirrors `snaq_flatnet` field for field
    n_nodes::Cint; n_edges::Cint; n_hybrids::Cint; n_leaves::Cint
    node_number::Ptr{Cint}; node_flags::Ptr{Cuint}; node_incycle::Ptr{Cint}; node_gammaz::Ptr{Cdouble}
    node_taxon::Ptr{Cint}; node_edge::Ptr{Cint}
    edge_number::Ptr{Cint}; edge_flags::Ptr{Cuint}; edge_incycle::Ptr{Cint}; edge_length::Ptr{Cdouble}
    edge_gamma::Ptr{Cdouble}; edge_node::Ptr{Cint}
    hybrid::Ptr{Cint}; leaf::Ptr{Cint}
end
struct Opts; device::Cint; rank::Cint; world_size::Cint; flags::Cuint; end     # flags: 2 = SNAQ_OPT_PER_QUARTET (default 0: run table)

mutable struct Engine
    ctx::Ptr{Cvoid}
    taxa::Vector{String}
    nq::Int
    last_x::Vector{Float64}      # parameters of the last optimisation (weights of the deltaCF-weighted quartet draw)
    has_net::Bool
end

check(e::Engine, rc) = rc == 0 || error(unsafe_string(ccall((:snaq_last_error, LIB), Cstring, (Ptr{Cvoid},), e.ctx)))

"one engine per Julia worker process; worker i uses GPU (i-1) % ngpus (pmap over runs: src/snaq_optimization.jl:1681)"
function Engine(d::DataCF; device::Integer = (Distributed.myid() - 1) % 8, per_quartet::Bool = false, rank::Integer = 0, world_size::Integer = 1)
    ctx = Ref{Ptr{Cvoid}}(C_NULL)   # per_quartet: the per-quartet kernels instead of the run table with run-summed weights (include/snaq_b200.h)
    rc = ccall((:snaq_create, LIB), Cint, (Ref{Opts}, Ref{Ptr{Cvoid}}), Opts(device, rank, world_size, per_quartet ? 2 : 0), ctx)
    rc == 0 || error(unsafe_string(ccall((:snaq_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
    taxa = SNaQ.tiplabels(d)                               # src/readquartetdata.jl:440
    tid = Dict(t => UInt16(i - 1) for (i, t) in enumerate(taxa))
    qt = Matrix{UInt16}(undef, 4, d.numQuartets); obs = Matrix{Float64}(undef, 3, d.numQuartets)
    for (j, q) in enumerate(d.quartet)
        for i in 1:4 qt[i, j] = tid[q.taxon[i]] end
        obs[:, j] = q.obsCF
    end
    e = Engine(ctx[], taxa, d.numQuartets, Float64[], false)
    check(e, ccall((:snaq_set_data, LIB), Cint, (Ptr{Cvoid}, Cint, Int64, Ptr{UInt16}, Ptr{Cdouble}),
                   e.ctx, length(taxa), d.numQuartets, qt, obs))
    finalizer(x -> ccall((:snaq_destroy, LIB), Cint, (Ptr{Cvoid},), x.ctx), e)
    return e
end

"serialise exactly the fields the hot path reads (SURVEY.md App. B) in net.node / net.edge / node.edge order;
returns the struct and the arrays it points into (keep them alive with GC.@preserve while the struct is in use)"
function flatnet(e::Engine, net::HybridNetwork)
    nidx = Dict(objectid(n) => Cint(i - 1) for (i, n) in enumerate(net.node))
    eidx = Dict(objectid(ed) => Cint(i - 1) for (i, ed) in enumerate(net.edge))
    tid = Dict(t => Cint(i - 1) for (i, t) in enumerate(e.taxa))
    nn, ne = length(net.node), length(net.edge)
    node_number = Cint[n.number for n in net.node]
    node_flags = Cuint[(n.leaf ? 0x01 : 0) | (n.hybrid ? 0x02 : 0) | (hasHybEdge(n) ? 0x04 : 0) | (isBadDiamondI(n) ? 0x08 : 0) |
                       (isBadDiamondII(n) ? 0x10 : 0) | (isBadTriangle(n) ? 0x20 : 0) for n in net.node]
    node_incycle = Cint[inCycle(n) for n in net.node]
    node_gammaz = Cdouble[gammaz(n) for n in net.node]
    node_taxon = Cint[n.leaf ? tid[n.name] : -1 for n in net.node]
    node_edge = fill(Cint(-1), 3, nn)
    for (i, n) in enumerate(net.node), (j, ed) in enumerate(n.edge) node_edge[j, i] = eidx[objectid(ed)] end
    edge_number = Cint[ed.number for ed in net.edge]
    edge_flags = Cuint[(ed.hybrid ? 0x01 : 0) | (ed.ismajor ? 0x02 : 0) | (ed.ischild1 ? 0x04 : 0) |
                       (istIdentifiable(ed) ? 0x08 : 0) | (fromBadDiamondI(ed) ? 0x10 : 0) for ed in net.edge]
    edge_incycle = Cint[inCycle(ed) for ed in net.edge]
    edge_length = Cdouble[ed.length for ed in net.edge]
    edge_gamma = Cdouble[ed.gamma for ed in net.edge]
    edge_node = Cint[nidx[objectid(ed.node[k])] for k in 1:2, ed in net.edge]
    hyb = Cint[nidx[objectid(n)] for n in net.hybrid]
    leaf = Cint[nidx[objectid(n)] for n in net.leaf]
    keep = (node_number, node_flags, node_incycle, node_gammaz, node_taxon, node_edge, edge_number, edge_flags, edge_incycle,
            edge_length, edge_gamma, edge_node, hyb, leaf)
    f = FlatNet(nn, ne, length(hyb), length(leaf), pointer(node_number), pointer(node_flags), pointer(node_incycle),
                pointer(node_gammaz), pointer(node_taxon), pointer(node_edge), pointer(edge_number), pointer(edge_flags),
                pointer(edge_incycle), pointer(edge_length), pointer(edge_gamma), pointer(edge_node), pointer(hyb), pointer(leaf))
    return f, keep
end

"extractQuartet!(net,d) on the device.  patch=true (the search loop edits ONE network in place, src/snaq_optimization.jl:1154-1198):
snaq_patch_network re-describes only the quartets with a taxon whose root path changed; the table is bit-identical to a rebuild"
function set_network!(e::Engine, net::HybridNetwork; patch::Bool = false)
    f, keep = flatnet(e, net)
    if patch
        GC.@preserve keep check(e, ccall((:snaq_patch_network, LIB), Cint, (Ptr{Cvoid}, Ref{FlatNet}, Ptr{Cint}, Cint), e.ctx, f, C_NULL, 0))
    else
        GC.@preserve keep check(e, ccall((:snaq_set_network, LIB), Cint, (Ptr{Cvoid}, Ref{FlatNet}), e.ctx, f))
    end
end

"objective value for one x (the body of `obj` in optBL!, src/snaq_optimization.jl:368-380)"
function pseudodeviance(e::Engine, x::Vector{Float64})
    out = Ref{Cdouble}(0.0)
    check(e, ccall((:snaq_eval, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cdouble}, Ref{Cdouble}), e.ctx, 1, length(x), x, out))
    return out[]
end

"batched: one row of X per candidate (used by a batched optimiser / speculative trust-region points)"
function pseudodeviance(e::Engine, X::Matrix{Float64})          # X is nparams x P (column-major = row-major P x nparams)
    out = Vector{Float64}(undef, size(X, 2))
    check(e, ccall((:snaq_eval, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}), e.ctx, size(X, 2), size(X, 1), X, out))
    return out
end

"q.qnet.expCF / q.logPseudoLik / q.deltaCF for fittedquartetCF, sampleEdgeQuartetWeighted (src/descriptive.jl:24-59)"
function readback!(e::Engine, d::DataCF, x::Vector{Float64})
    expcf = Matrix{Float64}(undef, 3, e.nq); lp = Vector{Float64}(undef, e.nq); dl = Vector{Float64}(undef, e.nq)
    check(e, ccall((:snaq_eval_quartets, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                   e.ctx, length(x), x, expcf, lp, dl))
    for (j, q) in enumerate(d.quartet)
        q.qnet.expCF = expcf[:, j]; q.logPseudoLik = lp[j]; q.deltaCF = dl[j]
    end
end

# ---- the replaced body of optBL! (src/snaq_optimization.jl:341-390): NLopt stays in Julia, only `obj` changes ----
# `readback`: copy expCF / logPseudoLik / deltaCF of every quartet into the Quartet objects afterwards, as the reference
# leaves them (156 bytes per quartet: 20 ms at 3.9 M quartets, several optBL! calls' worth).  Only the final report
# (fittedquartetCF, writeExpCF) needs them once the weighted draw of the proposals uses `sample_quartet`: off by default.
function optBL!(net::HybridNetwork, d::DataCF, e::Engine, verbose::Bool, ftolRel, ftolAbs, xtolRel, xtolAbs; readback::Bool = false)
    SNaQ.parameters!(net)                       # :353
    set_network!(e, net; patch = e.has_net)     # replaces extractQuartet!(net,d), :354 (incremental after the first topology)
    e.has_net = true
    nht = length(ht(net))
    opt = SNaQ.NLopt.Opt(:LN_BOBYQA, nht)
    SNaQ.NLopt.ftol_rel!(opt, ftolRel); SNaQ.NLopt.ftol_abs!(opt, ftolAbs)
    SNaQ.NLopt.xtol_rel!(opt, xtolRel); SNaQ.NLopt.xtol_abs!(opt, xtolAbs)
    SNaQ.NLopt.maxeval!(opt, 1000)
    SNaQ.NLopt.lower_bounds!(opt, zeros(nht)); SNaQ.NLopt.upper_bounds!(opt, SNaQ.upper(net))
    obj(x::Vector{Float64}, g::Vector{Float64}) = (SNaQ.update!(net, x); pseudodeviance(e, x))   # :368-380
    SNaQ.NLopt.min_objective!(opt, obj)
    fmin, xmin, ret = SNaQ.NLopt.optimize(opt, ht(net))
    SNaQ.updateParameters!(net, xmin)           # :387
    SNaQ.loglik!(net, fmin)                     # :388
    e.last_x = copy(xmin)
    readback && readback!(e, d, xmin)           # (quirk: the reference leaves the LAST x it tried there, not xmin)
end

# ---- phase 2: the optimiser itself on the device (snaq_optbl): analytic gradient in one pass, batched line searches ----
struct OptblOpts; ftol_rel::Cdouble; ftol_abs::Cdouble; xtol_rel::Cdouble; xtol_abs::Cdouble; maxeval::Cint; ladder::Cint; flags::Cuint; reserved::Cuint; end
mutable struct OptblResult; fmin::Cdouble; proj_grad::Cdouble; nevals::Cint; nlaunches::Cint; niter::Cint; status::Cint; end

"optBL! without NLopt: same arguments, same post-conditions (ht(net), loglik(net) updated); iterates differ from BOBYQA's"
function optBL_device!(net::HybridNetwork, d::DataCF, e::Engine, ftolRel = SNaQ.fRel, ftolAbs = SNaQ.fAbs, xtolRel = SNaQ.xRel, xtolAbs = SNaQ.xAbs;
                       readback::Bool = false)
    (ftolRel > 0 && ftolAbs > 0 && xtolAbs > 0 && xtolRel > 0) || error("tolerances have to be positive, ftol (rel,abs), xtol (rel,abs): $([ftolRel, ftolAbs, xtolRel, xtolAbs])")
    SNaQ.parameters!(net)
    set_network!(e, net; patch = e.has_net)
    e.has_net = true
    x = copy(ht(net))
    res = OptblResult(0, 0, 0, 0, 0, 0)
    check(e, ccall((:snaq_optbl, LIB), Cint, (Ptr{Cvoid}, Ref{OptblOpts}, Ptr{Cdouble}, Ref{OptblResult}), e.ctx,
                   OptblOpts(ftolRel, ftolAbs, xtolRel, xtolAbs, 1000, 0, 0, 0), x, res))
    SNaQ.updateParameters!(net, x)
    SNaQ.loglik!(net, res.fmin)
    e.last_x = copy(x)
    readback && readback!(e, d, x)
    return res
end

"optBL! of several proposed networks at once (snaq_optbl_topologies): the opt-in candidate batching of SURVEY 8f rank 2.
Each net gets what optBL_device! would have given it (ht, loglik); e's own network is left as it was."
function optBL_candidates!(nets::Vector{HybridNetwork}, d::DataCF, e::Engine; workers::Integer = 0,
                           ftolRel = SNaQ.fRel, ftolAbs = SNaQ.fAbs, xtolRel = SNaQ.xRel, xtolAbs = SNaQ.xAbs)
    foreach(SNaQ.parameters!, nets)
    flats = [flatnet(e, n) for n in nets]
    fs = FlatNet[f for (f, _) in flats]
    ldx = maximum(length(ht(n)) for n in nets; init = 1)
    xmin = zeros(Cdouble, ldx, length(nets))                  # column c = candidate c (row-major C x ldx on the C side)
    npar = zeros(Cint, length(nets))
    res = [OptblResult(0, 0, 0, 0, 0, 0) for _ in nets]
    raw = zeros(UInt8, 32 * length(nets))                     # OptblResult is mutable: a plain buffer, sizeof(snaq_optbl_result) = 2*8 + 4*4
    GC.@preserve flats check(e, ccall((:snaq_optbl_topologies, LIB), Cint,
        (Ptr{Cvoid}, Cint, Ptr{FlatNet}, Ref{OptblOpts}, Cint, Ptr{Cdouble}, Cint, Ptr{Cint}, Ptr{UInt8}),
        e.ctx, length(nets), fs, OptblOpts(ftolRel, ftolAbs, xtolRel, xtolAbs, 1000, 0, 0, 0), workers, xmin, ldx, npar, raw))
    for (c, net) in enumerate(nets)
        r = reinterpret(Float64, raw[32 * (c - 1) + 1:32 * (c - 1) + 16]); i = reinterpret(Int32, raw[32 * (c - 1) + 17:32 * c])
        res[c] = OptblResult(r[1], r[2], i[1], i[2], i[3], i[4])
        SNaQ.updateParameters!(net, xmin[1:npar[c], c])
        SNaQ.loglik!(net, res[c].fmin)
    end
    return res
end

"objective and gradient in one pass (for NLopt's :LD_LBFGS / :LD_MMA if NLopt is kept: fill g in the callback)"
function pseudodeviance_grad!(e::Engine, x::Vector{Float64}, g::Vector{Float64})
    f = Ref{Cdouble}(0.0)
    check(e, ccall((:snaq_eval_grad, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ref{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                   e.ctx, length(x), x, f, g, C_NULL))
    return f[]
end

"weighted draw of a quartet (weights deltaCF at x) for sampleEdgeQuartetWeighted, src/moves_snaq.jl:1443-1449: 1-based index"
function sample_quartet(e::Engine, x::Vector{Float64}, u::Float64 = rand())
    idx = Ref{Int64}(0)
    check(e, ccall((:snaq_sample_quartets, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Cint, Ref{Cdouble}, Ref{Int64}),
                   e.ctx, length(x), x, 1, u, idx))
    return idx[] + 1
end

"bootstrap replicate on the device: keep the CI columns resident once (set_ci!), then one call per replicate.
The draws come from the engine's counter-based generator, not from Julia's RNG: replicates differ from bootsnaq's for
the same seed (same distribution, src/bootstrap.jl:60-67)."
function set_ci!(e::Engine, lo::Matrix{Float64}, hi::Matrix{Float64})          # 3 x nq each (column-major = [nq][3])
    check(e, ccall((:snaq_set_ci, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}), e.ctx, lo, hi))
end
resample_obscf!(e::Engine, seed::Integer) =
    check(e, ccall((:snaq_resample_obscf, LIB), Cint, (Ptr{Cvoid}, UInt64), e.ctx, UInt64(seed)))

"updateBL! (src/readquartetdata.jl:838-861) with the table built on the device; net must be a tree"
function updateBL_device!(net::HybridNetwork, d::DataCF, e::Engine)
    SNaQ.parameters!(net)
    set_network!(e, net)
    nh = net.numhybrids - numBad(net)
    nt = count(istIdentifiable, net.edge)
    nq = Vector{Int64}(undef, nt); el = Vector{Float64}(undef, nt)
    check(e, ccall((:snaq_update_bl, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Cdouble}), e.ctx, nq, el))
    for i in 1:nt
        nq[i] > 0 || continue
        ed = net.edge[SNaQ.getIndexEdge(numht(net)[nh + i], net)]
        if ed.length < 0.0 || ed.length == 1.0                  # :850-853
            SNaQ.setLength!(ed, el[i] > 0 ? el[i] : 0.0)
        end
    end
    for ed in net.edge
        ed.length < 0.0 && SNaQ.setLength!(ed, 1.0)             # :855-858
    end
end

# ---- quartet-sharded mode: the collective lives in the library (snaq_comm_unique_id / snaq_comm_init) ----
"rank 0 creates the 128-byte id and hands it to the other ranks (Distributed.remotecall / a file / MPI.bcast)"
function comm_unique_id()
    id = zeros(UInt8, 128)
    ccall((:snaq_comm_unique_id, LIB), Cint, (Ptr{UInt8},), id) == 0 || error(unsafe_string(ccall((:snaq_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
    return id
end
"collective: every rank of a sharded engine (Engine(d; rank, world_size)) calls it with the same id; afterwards\npseudodeviance / pseudodeviance_grad! / optBL_device! return totals over all ranks"
comm_init!(e::Engine, id::Vector{UInt8}) = check(e, ccall((:snaq_comm_init, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}), e.ctx, id))

"edge numbers of d.quartet[idx].qnet.edge (1-based idx), for step 2 of sampleEdgeQuartetWeighted (src/moves_snaq.jl:1451-1453)"
function quartet_edges(e::Engine, idx::Integer)
    out = Vector{Cint}(undef, 4096); n = Ref{Cint}(0)
    check(e, ccall((:snaq_get_quartet_edges, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Cint}, Cint, Ref{Cint}), e.ctx, idx - 1, out, length(out), n))
    return Int.(out[1:n[]])
end

# ---- the drop-in hook ---------------------------------------------------------------------------------------------------
# snaq!, bootsnaq, topologymaxQpseudolik! and topologyQpseudolik! reach the hot path through SNaQ's OWN functions
# (optTopLevel! calls optBL! at src/snaq_optimization.jl:1346, :1387, :1435; afterOptBL! at :727; gammaZero! at :581).
# install!() replaces those methods (same signatures) so that nothing above them changes.  Engines are kept in a registry
# keyed by the DataCF object: one per worker process and data set, created on first use, data resident in HBM.
const ENGINES = IdDict{Any,Engine}()
const MODE = Ref(:nlopt)            # :nlopt  = phase 1: NLopt's BOBYQA in Julia, objective on the device (the reference's search path)
                                    # :device = phase 2: snaq_optbl (projected quasi-Newton on the analytic gradient)
engine_for(d::DataCF; kwargs...) = get!(() -> Engine(d; kwargs...), ENGINES, d)

"push q.sampled to the engine (writers: updateSubsetQuartets!, updateUninformativeQuartets!, src/update.jl:386,406,442)"
function sync_sampled!(e::Engine, d::DataCF)
    mask = UInt8[q.sampled ? 1 : 0 for q in d.quartet]
    check(e, ccall((:snaq_set_sampled, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}), e.ctx, mask))
end
"push the observed CFs again (bootstrap: readtableCF!(data, df, cols) overwrites q.obsCF in place, src/readquartetdata.jl:211-217)"
function sync_obscf!(e::Engine, d::DataCF)
    obs = Matrix{Float64}(undef, 3, d.numQuartets)
    for (j, q) in enumerate(d.quartet) obs[:, j] = q.obsCF end
    check(e, ccall((:snaq_set_obscf, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), e.ctx, obs))
end

function install!(; mode::Symbol = :nlopt)
    mode in (:nlopt, :device) || error("mode is :nlopt or :device")
    MODE[] = mode
    @eval SNaQ begin
        # optBL!: src/snaq_optimization.jl:341-390
        function optBL!(net::HybridNetwork, d::DataCF, verbose::Bool, ftolRel::Float64, ftolAbs::Float64, xtolRel::Float64, xtolAbs::Float64)
            (ftolRel > 0 && ftolAbs > 0 && xtolAbs > 0 && xtolRel > 0) || error("tolerances have to be positive, ftol (rel,abs), xtol (rel,abs): $([ftolRel, ftolAbs, xtolRel, xtolAbs])")
            e = $(engine_for)(d)
            if $(MODE)[] === :device
                $(optBL_device!)(net, d, e, ftolRel, ftolAbs, xtolRel, xtolAbs)
            else
                $(optBL!)(net, d, e, verbose, ftolRel, ftolAbs, xtolRel, xtolAbs)
            end
            return nothing
        end
        # topologyQpseudolik!: src/pseudolik.jl:1342-1378 (same preparation of the network; the quartet loop runs on the device)
        function topologyQpseudolik!(net0::HybridNetwork, d::DataCF; verbose = false::Bool)
            for ed in net0.edge
                !ed.hybrid || (ed.gamma >= 0.0) || error("hybrid edge has missing γ value. Cannot compute quartet pseudo-likelihood.\nTry `topologymaxQpseudolik!` instead, to estimate these γ's.")
            end
            net = readTopologyUpdate(writenewick_level1(net0))
            if !isempty(d.repSpecies)
                expandLeaves!(d.repSpecies, net)
                net = readnewicklevel1(writenewick_level1(net))
            end
            try checkNet(net) catch; error("starting topology not a level 1 network") end
            e = $(engine_for)(d)
            parameters!(net)
            $(set_network!)(e, net)
            val = $(pseudodeviance)(e, ht(net))
            $(readback!)(e, d, ht(net))                 # q.qnet.expCF for fittedquartetCF / writeExpCF
            loglik!(net0, val)
            return val
        end
        # sampleEdgeQuartetWeighted: src/moves_snaq.jl:1434-1468.  The weighted draw and the quartet's edge list come from the
        # engine; the RNG calls are the reference's (one rand() for the quartet, one sample() among its edges).
        function sampleEdgeQuartetWeighted(edges::Vector{Edge}, d::DataCF)
            e = $(engine_for)(d)
            x = $(current_x)(e)
            edge_numbers = [ed.number for ed in edges]
            while true
                idx = $(sample_quartet)(e, x, rand())
                q_edges = [n for n in $(quartet_edges)(e, idx) if n in edge_numbers]
                isempty(q_edges) && continue
                edge = sample(q_edges, 1)[1]
                return findfirst(isequal(edge), q_edges)
            end
        end
        # data refresh hooks: the engine's copy follows the DataCF object
        function readtableCF!(data::DataCF, df::DataFrames.DataFrame, cols::Vector{<:Integer})
            for i in 1:size(df, 1), j in 1:3
                data.quartet[i].obsCF[j] = df[i, cols[j]]
            end
            haskey($(ENGINES), data) && $(sync_obscf!)($(ENGINES)[data], data)
        end
    end
    return nothing
end

"the parameter vector of the last optBL! on this engine (the weights of the next proposal's quartet draw are deltaCF there)"
current_x(e::Engine) = e.last_x

end # module
