# AKSB200.jl -- ccall shim: AtomicKohnSham.jl's SCF loop on libaks_b200.so (include/aks_b200.h).
#
# NOT EXECUTABLE IN THIS REPOSITORY'S ENVIRONMENT (no Julia in the image or on the GPU box); it is the
# reference-side binding a maintainer would add.  It keeps the public API: the caller builds
# KSEModel / KSEDiscretization / ODA exactly as today and passes `backend = :b200`.
module AKSB200

using AtomicKohnSham
using AtomicKohnSham: KSESolver, KSESolution, init_cache!, register!, callback!, findindex
using LinearAlgebra, SparseArrays

const LIB = get(ENV, "AKS_B200_LIB", "libaks_b200.so")

struct IterRecord              # aks_iter_record
    niter::Int32; status::Int32
    stop::Float64; Etot::Float64; Ekin::Float64; Ecou::Float64; Ehar::Float64; Eexc::Float64; t::Float64
end

struct DiscDesc                # aks_disc_desc (field order must match the header)
    dtype::Int32; p::Int32; C::Int32; Nh::Int32; Q::Int32; G::Int32; lh::Int32; nh::Int32; nspin::Int32
    mesh::Ptr{Float64}; invshifts::Ptr{Float64}; x::Ptr{Float64}; w::Ptr{Float64}; Pgenx::Ptr{Float64}
    y::Ptr{Float64}; wy::Ptr{Float64}; ycell::Ptr{Int64}; Pgeny::Ptr{Float64}
    M0::Ptr{Float64}; A::Ptr{Float64}; Mm1::Ptr{Float64}; Mm2::Ptr{Float64}
    nF::Int64; Fi::Ptr{Int64}; Fj::Ptr{Int64}; Fm::Ptr{Int64}; Fv::Ptr{Float64}
    cell_len::Ptr{Int64}; cell_indices::Ptr{Int64}; cell_generators::Ptr{Int64}
end

check(ctx, rc) = rc == 0 || rc == -3 ||   # AKS_ENOTCONV is reported through the status words
    error("aks_b200: status $rc: " * unsafe_string(ccall((:aks_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx)))

xc_id(f) = f isa AtomicKohnSham.NoFunctional ? Int32(0) :
           f.identifier == :lda_x ? Int32(1) : f.identifier == :lda_c_pw ? Int32(12) :
           error("backend :b200 supports lda_x / lda_c_pw only")

"""
    groundstate_b200(models, disc, alg; maxiter = 100, callback = nothing, device = 0)

Drop-in for `groundstate` on one model or a vector of models sharing `disc`.  The tables are the ones
`init_cache!` / `GaussLegendre` build on the host; every SCF iteration is one `aks_scf_step`.
"""
function groundstate_b200(models::Vector{<:KSEModel}, disc, algs; maxiter = 100, callback = nothing, device = 0,
                          plot_grid = nothing)
    # `algs` / `maxiter`: one ODA / Int for every model, or a vector with one entry per model (the reference picks
    # both per atom: examples/atomic density comparison/run.jl:74-76) -> aks_batch_set_oda_problem
    alg = algs isa AbstractVector ? first(algs) : algs
    maxit_of(i) = maxiter isa AbstractVector ? maxiter[i] : maxiter
    maxit_all = maxiter isa AbstractVector ? maximum(maxiter) : maxiter
    basis, q = disc.basis, disc.fem_integration_method
    init_cache!(disc, first(models))                      # A, M0, M-1, M-2, F (host, once)
    ops = disc.femops
    p = basis.generators.ordermax; C = length(basis.mesh) - 1; n = p + 1
    # tables the reference recomputes on the fly, built once here
    ycell = Int64[findindex(basis.mesh, yy) for yy in q.y]
    Pgeny = hcat((begin ϕ = basis.shifts[c]; AtomicKohnSham.evaluate(basis.generators.polynomials, ϕ[1]*yy + ϕ[2]) end
                  for (c, yy) in zip(ycell, q.y))...)
    cell_len = Int64[length(v) for v in basis.cells_to_indices]
    cidx = zeros(Int64, n, C); cgen = zeros(Int64, n, C)   # column c = cell c  (== row-major [C][p+1] in C)
    for c in 1:C
        cidx[1:cell_len[c], c] .= basis.cells_to_indices[c]; cgen[1:cell_len[c], c] .= basis.cells_to_generators[c]
    end
    # T = Float64 -> dtype 0.  T = Double64 -> dtype 1: a Vector{Double64} IS an array of interleaved (hi, lo)
    # Float64 pairs, so every table is handed over as it lies in memory (fp(v) below only retypes the pointer).
    T = eltype(basis)
    dtype = T === Float64 ? Int32(0) : Int32(1)
    fp(v) = Ptr{Float64}(pointer(v))
    invsh = T[v for s in basis.invshifts for v in s]
    Fk = collect(keys(ops.F)); Fi = Int64[k[1] for k in Fk]; Fj = Int64[k[2] for k in Fk]; Fm = Int64[k[3] for k in Fk]
    Fv = T[ops.F[k] for k in Fk]
    A = Matrix(ops.A); Mm1 = Matrix(ops.M₋₁); Mm2 = Matrix(ops.M₋₂); mesh = collect(T, basis.mesh.points)
    M0 = Matrix(ops.M₀); Pgeny = Matrix{T}(Pgeny)
    ctx = Ref{Ptr{Cvoid}}(); dh = Ref{Ptr{Cvoid}}(); bh = Ref{Ptr{Cvoid}}()
    check(C_NULL, ccall((:aks_ctx_create, LIB), Cint, (Cint, Ref{Ptr{Cvoid}}), device, ctx))
    GC.@preserve mesh invsh q ops M0 A Mm1 Mm2 Fi Fj Fm Fv cell_len cidx cgen ycell Pgeny begin
        desc = DiscDesc(dtype, p, C, disc.Nₕ, q.npoints, q.npoints, disc.lₕ, disc.nₕ, disc.nspin,
            fp(mesh), fp(invsh), fp(q.x), fp(q.w), fp(q.Pgenx), fp(q.y), fp(q.wy),
            pointer(ycell), fp(Pgeny), fp(M0), fp(A), fp(Mm1), fp(Mm2),
            length(Fv), pointer(Fi), pointer(Fj), pointer(Fm), fp(Fv), pointer(cell_len), pointer(cidx), pointer(cgen))
        check(ctx[], ccall((:aks_disc_create, LIB), Cint, (Ptr{Cvoid}, Ref{DiscDesc}, Ref{Ptr{Cvoid}}), ctx[], desc, dh))
    end
    nprob = length(models)
    Z = Float64[m.Z for m in models]; N = Float64[m.N for m in models]; har = Float64[m.hartree for m in models]
    ex = Int32[xc_id(m.exchange) for m in models]; ec = Int32[xc_id(m.correlation) for m in models]
    check(ctx[], ccall((:aks_batch_create, LIB), Cint,
        (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ref{Ptr{Cvoid}}),
        dh[], nprob, Z, N, har, ex, ec, C_NULL, C_NULL, bh))
    a = alg.aufbau
    opt = a isa OptimizedAufbau ? a : OptimizedAufbau()
    check(ctx[], ccall((:aks_batch_set_oda, LIB), Cint,
        (Ptr{Cvoid}, Float64, Float64, Int32, Int32, Float64, Float64, Int32, Float64, Int32),
        bh[], Float64(alg.scftol), Float64(alg.tinit), alg.frozen_t, alg.maxiter_ls, Float64(alg.abstol_ls),
        Float64(alg.reltol_ls), opt.max_degen, Float64(opt.tol), opt.handle_degen_spin_1iter))
    if a isa AtomicKohnSham.SmearedAufbau
        check(ctx[], ccall((:aks_batch_set_aufbau, LIB), Cint, (Ptr{Cvoid}, Int32, Float64, Ptr{Float64}),
            bh[], 1, Float64(a.Temp), C_NULL))
    elseif a isa AtomicKohnSham.FrozenAufbau
        # the FrozenAufbauCache holds nfix in the layout of n: (L, nh[, 2]); one copy per problem
        nfix = reduce(vcat, [vec(Float64.(AtomicKohnSham.create_cache_aufbau(disc, m, a).nfix)) for m in models])
        GC.@preserve nfix check(ctx[], ccall((:aks_batch_set_aufbau, LIB), Cint,
            (Ptr{Cvoid}, Int32, Float64, Ptr{Float64}), bh[], 2, 0.0, nfix))
    end
    if algs isa AbstractVector || maxiter isa AbstractVector
        for i in 1:nprob
            ai = algs isa AbstractVector ? algs[i] : algs
            check(ctx[], ccall((:aks_batch_set_oda_problem, LIB), Cint, (Ptr{Cvoid}, Int32, Float64, Float64, Int32, Int32),
                bh[], i - 1, Float64(ai.scftol), Float64(ai.tinit), ai.frozen_t, maxiter isa AbstractVector ? maxit_of(i) : 0))
        end
    end
    solvers = [KSESolver(m, disc, algs isa AbstractVector ? algs[i] : algs; maxiter = maxit_of(i), callback = callback)
               for (i, m) in enumerate(models)]
    recs = Vector{IterRecord}(undef, nprob)
    recs_dd = Array{T}(undef, 7, nprob)     # Double64 runs: stop, Etot, Ekin, Ecou, Ehar, Eexc, t with all digits
    if callback === nothing && dtype == 0
        # solve! in ONE ccall: convergence is decided on the device, the per-iteration records LogBook wants are
        # written there and copied once (aks_scf_run); register! is replayed from the history
        hist = Matrix{IterRecord}(undef, nprob, maxit_all)
        rc = ccall((:aks_scf_run, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{IterRecord}, Ptr{IterRecord}), bh[], maxit_all, hist, recs)
        rc in (0, -3) || check(ctx[], rc)                     # -3 = AKS_ENOTCONV: some problem ended with MAXITERS
        for (i, s) in enumerate(solvers), it in 1:recs[i].niter
            r = hist[i, it]; e = s.energies
            s.stopping_criteria = r.stop
            e.Etot, e.Ekin, e.Ecou, e.Ehar, e.Eexc = r.Etot, r.Ekin, r.Ecou, r.Ehar, r.Eexc
            s.algcache.t = r.t
            register!(s); s.niter += 1
        end
    else
    # solve!: one ccall per iteration so that register!/callback! fire as in groundstate.jl:55-57
    for it in 1:maxit_all
        check(ctx[], ccall((:aks_scf_step, LIB), Cint, (Ptr{Cvoid}, Ptr{IterRecord}), bh[], recs))
        dtype == 1 && check(ctx[], ccall((:aks_get_records_dd, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), bh[], fp(recs_dd)))
        running = false
        for (i, (s, r)) in enumerate(zip(solvers, recs))
            r.niter == it || continue                       # this problem had already stopped
            e = s.energies
            if dtype == 1
                s.stopping_criteria = recs_dd[1, i]
                e.Etot, e.Ekin, e.Ecou, e.Ehar, e.Eexc = recs_dd[2, i], recs_dd[3, i], recs_dd[4, i], recs_dd[5, i], recs_dd[6, i]
                s.algcache.t = recs_dd[7, i]
            else
                s.stopping_criteria = r.stop
                e.Etot, e.Ekin, e.Ecou, e.Ehar, e.Eexc = r.Etot, r.Ekin, r.Ecou, r.Ehar, r.Eexc
                s.algcache.t = r.t
            end
            register!(s); callback!(s.callback, s); s.niter += 1
            running |= r.status == 0
        end
        running || break
    end
    end
    # copy D, U, ϵ, n, W back into the ODACache so that KSESolution(solver, name) works untouched
    sols = map(enumerate(solvers)) do (i, s)
        c = s.algcache
        W = disc.cache.hartw.W
        check(ctx[], ccall((:aks_get_state, LIB), Cint,
            (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
            bh[], i - 1, fp(c.D), fp(c.U), fp(c.ϵ), fp(c.n), isempty(W) ? C_NULL : fp(W)))   # (hi, lo) pairs when T = Double64
        KSESolution(s, _solution_name(models[i]))
    end
    # optional: the post-processing curves on a plot grid straight from the device state (aks_eval_grid: eval_density,
    # eval_hartree, eval_vxc of src/solution/postprocess.jl:81-352) -- returned next to the solutions as (ρ, V_H, v_xc) per
    # problem; the host functions of postprocess.jl keep working on the KSESolution objects above.
    curves = nothing
    if plot_grid !== nothing && dtype == 0
        X = Vector{Float64}(plot_grid)
        evalg(i, kind, σ) = (out = similar(X); check(ctx[], ccall((:aks_eval_grid, LIB), Cint,
            (Ptr{Cvoid}, Int32, Int32, Int32, Int32, Int32, Int64, Ptr{Float64}, Ptr{Float64}),
            bh[], i - 1, kind, 0, 0, σ, length(X), X, out)); out)
        curves = [(ρ = evalg(i, 1, -1), VH = evalg(i, 2, 0), vxc = evalg(i, 3, 0)) for i in eachindex(models)]
    end
    ccall((:aks_batch_destroy, LIB), Cvoid, (Ptr{Cvoid},), bh[])
    ccall((:aks_disc_destroy, LIB), Cvoid, (Ptr{Cvoid},), dh[])
    ccall((:aks_ctx_destroy, LIB), Cvoid, (Ptr{Cvoid},), ctx[])
    curves === nothing ? sols : (sols, curves)
end

# Double64 tables built by the library on the device instead of QuadGK.gauss(Double64, n) + init_cache! on the host
# (src/fem/integration methods.jl:46-70, src/discretization/discretization.jl:320-341); results are (hi, lo) pairs,
# i.e. exactly the memory of a Vector{Double64}.
function gauss_legendre_b200(n::Integer; device = 0)
    ctx = Ref{Ptr{Cvoid}}(C_NULL)
    check(C_NULL, ccall((:aks_ctx_create, LIB), Cint, (Cint, Ref{Ptr{Cvoid}}), device, ctx))
    x = Vector{Double64}(undef, n); w = Vector{Double64}(undef, n)
    GC.@preserve x w check(ctx[], ccall((:aks_dd_gauss_legendre, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}),
                                        ctx[], n, Ptr{Float64}(pointer(x)), Ptr{Float64}(pointer(w))))
    ccall((:aks_ctx_destroy, LIB), Cvoid, (Ptr{Cvoid},), ctx[])
    x, w
end

groundstate_b200(model::KSEModel, disc, alg; kw...) = only(groundstate_b200([model], disc, alg; kw...))


# same naming as groundstate (src/solver/groundstate.jl:24-34)
function _solution_name(model)
    Z, N = model.Z, model.N
    floor(Z) == Z || return "Z=$Z, N=$N"
    sym = AtomicKohnSham.ATOMIC_NUMBER_TO_NAME[Int(Z)]
    dq = Z - N
    iszero(dq) ? sym : sym * string(abs(dq)) * (dq > 0 ? "+" : "-")
end

end # module
