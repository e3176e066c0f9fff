# ILAB200Overloads.jl — the method overloads that keep InfiniteLinearAlgebra.jl's own API a drop-in over libila_b200.so.
#
# NOT EXECUTED in this repository (no Julia in the build image nor on the GPU box, SURVEY §0.2): this is the patch a
# maintainer applies.  `include` it from src/InfiniteLinearAlgebra.jl after the stock files (it REPLACES the bodies of
# the methods named below, file:line given for each); ILAB200.jl (same directory) holds the `ccall`s.
#
#   ql(A - λ*I), ql(A; branch)        -> ql_hessenberg(::InfToeplitz)        src/banded/infqltoeplitz.jl:56-65
#   ql(A::PertToeplitz)               -> ql_hessenberg!(B) tail step         src/infql.jl:70-76 (through the method above)
#   ql_L11(A, λs), ℓ11 over a grid    -> new batched forms                    examples/toeplitz.jl:14-33
#   qr(A) \ b, A \ b                  -> \(F::QR{…AdaptiveQRFactors}, b)     src/infqr.jl:370-374
#   partialqr!(F, n) (batch of one)   -> resumable device state               src/infqr.jl:32-48
#   cholesky(S) \ b                   -> \(F::Cholesky{…AdaptiveCholeskyFactors}, b)   src/infcholesky.jl:71-75, 94-140
#   reversecholesky(::SymTridiagonal) leading block, TridiagonalConjugation resizedata!   (consumers, SURVEY §8f)
#
# Everything the library cannot describe (non-affine bands, u ≥ 2 symbols, custom `branch` selectors, exotic eltypes)
# falls through to the stock Julia method with `invoke`, so behaviour outside the accelerated path is unchanged.

import .ILAB200
import LinearAlgebra: ldiv!, QR, Cholesky, UpperTriangular
import MatrixFactorizations: QL, ql
import BandedMatrices: _BandedMatrix, bandwidths, BandedMatrix
import FillArrays: Fill, Ones, Zeros, AbstractFill, getindex_value
import LazyArrays: Vcat, Hcat, ApplyArray, arguments, paddeddata
import InfiniteArrays: ℵ₀, OneToInf, InfUnitRange, InfStepRange, ∞

# ======================================================================================================================
# 1. Toeplitz QL: ql(A - λ*I) for A::InfToeplitz with u == 1
# ======================================================================================================================

# `branch` of the reference is any function (n2) -> (value, index) (src/banded/infqltoeplitz.jl:7,18; default findmax; the
# examples pass k-th-largest selectors, examples/toeplitz.jl:22-27).  The library takes the RANK k.
struct BranchRank
    k::Int
end
(b::BranchRank)(n2) = (s = sortperm(n2); (n2[s[end-b.k+1]], s[end-b.k+1]))
branch_rank(::typeof(findmax)) = 1
branch_rank(b::BranchRank) = b.k
branch_rank(_) = nothing                      # unknown selector: stock Julia path

# REPLACES src/banded/infqltoeplitz.jl:56-65.  Same result object, built from the library's record exactly as the
# reference builds it from `tail_de` + `ql_X!` (:63-64).
function ql_hessenberg(A::InfToeplitz{T}; branch=findmax, kwds...) where {T<:Union{Float64,ComplexF64}}
    l, u = bandwidths(A)
    @assert u == 1
    k = branch_rank(branch)
    (k === nothing || l > 7) && return invoke(ql_hessenberg, Tuple{InfToeplitz}, A; branch=branch, kwds...)
    col = A.data.args[1]                      # [a₊₁, a₀, a₋₁, …, a₋ₗ]; the library reverses it (:59)
    Xs, τs, st = ILAB200.ql_tail(col, [zero(T)]; branch_rank=k)
    ILAB200.throw_status(st[1], reverse(col))                      # DomainError of tail_de (:12-13, :23)
    X = Xs[:, :, 1]::Matrix{T}
    F = QL(X, τs[:, 1])                                            # what ql_X! returns: factors X, τ
    factors = _BandedMatrix(Hcat([zero(T); X[1, end-1]; X[2, end-1:-1:1]], [0; X[2, end:-1:1]] * Ones{T}(1, ∞)), ℵ₀, l + u, 1)
    QLHessenberg(factors, Fill(F.Q, ∞))
end

# The batched form over a vector of shifts (new API, north star): L[1,1] of ql(A - λ*I) for every λ.
"""
    ql_L11(A::InfToeplitz, λs; branch=findmax, ngpus=1) -> Vector

`[ql(A - λ*I).L[1,1] for λ in λs]` in one library call; throws the reference's `DomainError` at the first shift without
a QL factorization unless `throw=false`, in which case those entries are `NaN` and `status` is returned as well.
"""
function ql_L11(A::InfToeplitz{T}, λs::AbstractVector; branch=findmax, ngpus::Integer=1, throw::Bool=true) where {T}
    k = something(branch_rank(branch), 0)
    k == 0 && return [ql(A - λ * I; branch=branch).L[1, 1] for λ in λs]
    col = A.data.args[1]
    if T <: Real && eltype(λs) <: Real
        Xs, _, st = ILAB200.ql_tail(col, λs; branch_rank=k, ngpus=ngpus)
        L = [Xs[1, end-1, i] for i in eachindex(λs)]               # L[1,1] = X[1, m-1] (src/banded/infqltoeplitz.jl:63)
    else
        Lre, st = ILAB200.ql_L11_real(col, λs; branch_rank=k, ngpus=ngpus)
        L = complex.(Lre)                                          # eltype of the reference's result
    end
    if throw
        i = findfirst(!iszero, st)
        i === nothing || ILAB200.throw_status(st[i], λs[i])
        return L
    end
    L, st
end

# examples/toeplitz.jl:14-33 — `ℓ11.(Ref(A), x' .+ y.*im)` becomes one call; -1.0 where no QL exists (the example's catch)
ℓ11_grid(A::InfToeplitz, x::AbstractVector{<:Real}, y::AbstractVector{<:Real}; branch=findmax, ngpus::Integer=1) =
    ILAB200.absl11_grid(A.data.args[1], x, y; branch_rank=something(branch_rank(branch), 1), ngpus=ngpus)

# `ql(A - λI) \ b` over many shifts (resolvent / spectral-measure loops): first `nout` entries per shift, n × nout
ql_solve(A::InfToeplitz, λs::AbstractVector, b::AbstractVector, nout::Integer; kw...) =
    ILAB200.ql_solve(A.data.args[1], λs, b, nout; kw...)

# ======================================================================================================================
# 2. Operator description for the adaptive QR / Cholesky entries: per-diagonal generators (p + q t)/r (+ head, − shift)
# ======================================================================================================================

# One band of an ∞ BandedMatrix as the library's generator triple, or `nothing` if it is not affine in the position t.
# Entries must come out BIT-identical to what getindex on the lazy band gives, hence the dispatch on the lazy types:
#   Fill / Ones / Zeros        value                      -> (v, 0, 1)
#   InfUnitRange a:∞           a + t                      -> (a, 1, 1)
#   InfStepRange a:s:±∞        a + t*s (Julia's formula)  -> (a, s, 1)
#   Vcat(head::Vector, tail)   explicit head, then a generator shifted by length(head) positions
generator(v::AbstractFill) = (Float64(getindex_value(v)), 0.0, 1.0, Float64[])
generator(v::InfUnitRange) = (Float64(first(v)), 1.0, 1.0, Float64[])
generator(v::InfStepRange) = (Float64(first(v)), Float64(step(v)), 1.0, Float64[])
function generator(v::ApplyArray{<:Any,1,typeof(vcat)})
    args = arguments(v)
    length(args) == 2 && args[1] isa AbstractVector && (g = generator(args[2])) !== nothing || return nothing
    h = Vector{Float64}(args[1])
    p, q, r, _ = g
    (p - q * length(h), q, r, h)              # position t of the band = length(h) + (position in the tail)
end
generator(_) = nothing

"`(l, u, p, q, r, head)` describing `A` for `ILAB200.qr_solve`, or `nothing` when some band is not affine."
function affine_description(A::BandedMatrix{T,<:Any,OneToInf{Int}}) where {T<:Real}
    l, u = bandwidths(A)
    rows = A.data isa ApplyArray{<:Any,2,typeof(vcat)} ? arguments(A.data) : nothing
    rows === nothing && return nothing
    gens = map(row -> generator(vec(row)), rows)                   # data row rr <-> diagonal u - rr, as the ABI orders them
    any(isnothing, gens) && return nothing
    hp = maximum(g -> length(g[4]), gens)
    # data row rr starts at column max(1, 1 + (u - rr)): BandedMatrices stores band d = u - rr shifted right by d columns,
    # so position t along the diagonal = column index - 1 - max(d, 0); the library wants generators in t
    p = Float64[]; q = Float64[]; r = Float64[]; head = zeros(length(gens), hp)
    for (rr, (gp, gq, gr, h)) in enumerate(gens)
        d = u - (rr - 1)
        off = max(d, 0)                                            # first stored entry of this data row sits at column 1 + off
        push!(p, gp + gq * off); push!(q, gq); push!(r, gr)
        for t in 1:hp
            head[rr, t] = t + off <= length(h) ? h[t+off] : (gp + gq * (t - 1 + off)) / gr
        end
    end
    (l, u, reshape(p, 1, :), reshape(q, 1, :), reshape(r, 1, :), hp == 0 ? nothing : head)
end
affine_description(_) = nothing

# ======================================================================================================================
# 3. Adaptive QR: qr(A) \ b, A \ b
# ======================================================================================================================

# REPLACES src/infqr.jl:373-374 (`\(F::QR{<:Any,<:AdaptiveQRFactors}, B::AbstractVector; kwds...) = ldiv!(F.R, F.Q'*B)`).
# The factorization object stays what `qr(A)` returns (lazy, stock); the SOLVE is one fused library call — sweep, adaptive
# Q'b with the COLGROWTH cadence, back-substitution — when A's bands are affine; otherwise the stock two-step path.
function \(F::QR{<:Any,<:AdaptiveQRFactors{Float64}}, B::AbstractVector; tolerance=floatmin(Float64), ngpus::Integer=1, ncap=nothing)
    A = F.factors.data.data.array                                   # the operator behind the CachedArray (src/infqr.jl:12)
    desc = affine_description(A)
    bd = B isa AbstractVector{<:Real} ? paddeddata(B) : nothing    # finitely supported right-hand side
    if desc === nothing || bd === nothing || desc[1] < 1
        return ldiv!(F.R, F.Q' * B)                                # stock path (src/infqr.jl:373)
    end
    l, u, p, q, r, head = desc
    cap = something(ncap, 200_000)
    while true
        x, ncols, nconv, st = ILAB200.qr_solve(l, u, p, q, r, reshape(Vector{Float64}(bd), 1, :); head=head, tol=tolerance,
                                               ncap=cap, ngpus=ngpus)
        st[1] == 3 && ncap === nothing && cap < 50_000_000 && (cap *= 4; continue)     # the reference's cache just keeps growing
        ILAB200.throw_status(st[1], "adaptive QR")
        return Vcat(x[1], Zeros{Float64}(∞))                       # same lazy shape the reference returns (datasize prefix + zeros)
    end
end

# Batched form: one operator family, many shifts / parameters.  `ops` is anything `affine_description` understands.
"""
    qr_solve_batch(As::AbstractVector{<:BandedMatrix}, bs; tolerance=floatmin(Float64), ncap, ngpus=1)

`[A \\ b for (A, b) in zip(As, bs)]` through one library call (all operators must share their bandwidths).
"""
function qr_solve_batch(As::AbstractVector, bs::AbstractVector; tolerance=floatmin(Float64), ncap::Integer, ngpus::Integer=1)
    ds = map(affine_description, As)
    any(isnothing, ds) && return [A \ b for (A, b) in zip(As, bs)]
    l, u = ds[1][1], ds[1][2]
    all(d -> d[1] == l && d[2] == u && d[6] === nothing, ds) || return [A \ b for (A, b) in zip(As, bs)]
    cat(k) = reduce(vcat, (d[k] for d in ds))
    nb = maximum(length ∘ paddeddata, bs)
    b = zeros(length(As), nb)
    for (i, bi) in enumerate(bs)
        d = paddeddata(bi); b[i, 1:length(d)] .= d
    end
    x, ncols, nconv, st = ILAB200.qr_solve(l, u, cat(3), cat(4), cat(5), b; tol=tolerance, ncap=ncap, ngpus=ngpus)
    foreach(i -> ILAB200.throw_status(st[i], "adaptive QR, operator $i"), eachindex(st))
    [Vcat(xi, Zeros{Float64}(∞)) for xi in x]
end

# ======================================================================================================================
# 4. Adaptive Cholesky: cholesky(S) \ b
# ======================================================================================================================

# The hooks `ArrayLayouts._cholesky(::SymTridiagonalLayout / ::SymmetricLayout{<:AbstractBandedLayout}, …)` stay as they are
# (src/infcholesky.jl:74-75: they return the lazy `Cholesky(AdaptiveCholeskyFactors(A), :U, 0)`).  What is replaced is the
# solve — `ldiv!(U, ldiv!(U', b))` = the adaptive forward substitution (:94-133) and U \ y (:135-140):
function \(F::Cholesky{Float64,<:AdaptiveCholeskyFactors{Float64}}, B::AbstractVector{<:Real}; tolerance=floatmin(Float64),
           ngpus::Integer=1, ncap=nothing)
    S = F.factors.data.array                                        # the symmetric operator behind the cache (src/infcholesky.jl:13-16)
    desc = affine_description(S isa BandedMatrix ? S : BandedMatrix(S))
    bd = paddeddata(B)
    desc === nothing && return ldiv!(F.U, ldiv!(F.U', copy(B)))    # stock path
    l, u, p, q, r, head = desc
    up = 1:u+1                                                      # upper bands only: data rows +u … 0
    cap = something(ncap, 200_000)
    while true
        x, ncols, nconv, st = ILAB200.chol_solve(u, p[:, up], q[:, up], r[:, up], reshape(Vector{Float64}(bd), 1, :);
                                                 head=head === nothing ? nothing : head[up, :], tol=tolerance, ncap=cap, ngpus=ngpus)
        st[1] == 3 && ncap === nothing && cap < 50_000_000 && (cap *= 4; continue)
        ILAB200.throw_status(st[1], "adaptive Cholesky")           # 4 -> PosDefException (upstream silently yields NaNs, :49)
        return Vcat(x[1], Zeros{Float64}(∞))
    end
end

# ======================================================================================================================
# 5. Block-tridiagonal perturbed Toeplitz: ql(A - x*I).L[1,1] over energies (examples/periodicschrodinger.jl:13-27)
# ======================================================================================================================
function ql_L11(A::BlockTriPertToeplitz{Float64}, energies::AbstractVector{<:Real}; atol=1e-10, maxit=1_000_000, ngpus::Integer=1)
    N = max(length(A.blocks.dl.args[1]), length(A.blocks.d.args[1]), length(A.blocks.du.args[1])) + 1     # src/infql.jl:190
    c, a, b = A[Block(N + 1, N)], A[Block(N, N)], A[Block(N - 1, N)]
    heads(v) = [Matrix{Float64}(m) for m in v.args[1]]
    L, iters, st = ILAB200.blocktri_L11(heads(A.blocks.dl), heads(A.blocks.d), heads(A.blocks.du), Matrix(c), Matrix(a), Matrix(b),
                                        energies; atol=atol, maxit=maxit, ngpus=ngpus)
    foreach(i -> ILAB200.throw_status(st[i], energies[i]), eachindex(st))          # 3 -> "Did not converge" (src/infql.jl:181)
    L
end

# ======================================================================================================================
# 6. Consumers: TridiagonalConjugation fills its bands through one call (src/banded/tridiagonalconjugation.jl:260-282)
# ======================================================================================================================
function resizedata!(data::TridiagonalConjugationData, n)
    n ≤ data.datasize && return data
    U = reduce(vcat, (permutedims(Vector{Float64}(data.U[band(k)][1:n+1])) for k in 0:bandwidth(data.U, 2)))
    X = vcat(permutedims(Vector{Float64}(data.X[band(-1)][1:n+1])), permutedims(Vector{Float64}(data.X[band(0)][1:n+1])),
             permutedims(Vector{Float64}(data.X[band(1)][1:n+1])))
    Y = ILAB200.triconj(U, X, n)                                   # 3 × n: dl, d, du — bit-identical to the two sequential sweeps
    resize!(data.Y.dl, n); resize!(data.Y.d, n); resize!(data.Y.du, n)
    copyto!(data.Y.dl, 1, view(Y, 1, :), 1, n); copyto!(data.Y.d, 1, view(Y, 2, :), 1, n); copyto!(data.Y.du, 1, view(Y, 3, :), 1, n)
    data.datasize = n
    data
end
