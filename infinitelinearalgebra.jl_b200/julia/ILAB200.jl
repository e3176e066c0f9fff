# ILAB200.jl — ccall shim over libila_b200.so (include/ila_b200.h).
#
# NOT EXECUTED in this repository's CI: Julia is not installed in the build image nor on the GPU box (SURVEY §0.2).
# It documents the binding a maintainer adds to InfiniteLinearAlgebra.jl so that the existing API stays a drop-in:
#   ql(A - λ*I).L[1,1]            -> ILAB200.ql_L11(A, [λ])[1]
#   ℓ11.(Ref(A), x' .+ y.*im)     -> ILAB200.ql_L11(A, vec(x' .+ y.*im))        (examples/toeplitz.jl:14-33)
#   A \ b  (∞ banded, affine bands) -> ILAB200.qr_solve(...)                     (src/infqr.jl:370-374)
#   ql(A - λ*I) \ b                -> ILAB200.ql_solve(A, λs, b, nout)           (src/infql.jl:266-289)
#   ql(::PertToeplitz), ql(::BlockTriPertToeplitz) -> ql_pert_L11, blocktri_L11   (src/infql.jl:70-93,189-209)
# Exceptions of the reference are rebuilt from status[] (SURVEY §5).
#
# This file binds EVERY entry of include/ila_b200.h that a Julia caller needs (batched forms on plain arrays).  The
# method overloads that make `ql(A - λ*I)`, `qr(A) \ b`, `cholesky(S) \ b` themselves route here live in
# ILAB200Overloads.jl next to it.
module ILAB200

using LinearAlgebra

const lib = get(ENV, "ILA_B200_LIB", "libila_b200.so")

struct IlaError <: Exception
    code::Cint
    msg::String
end
lasterror() = unsafe_string(ccall((:ila_last_error, lib), Cstring, ()))
check(rc) = rc == 0 || throw(IlaError(rc, lasterror()))

"""
    PinnedVector{T}(n) -> Vector{T}

A `Vector{T}` on page-locked host memory from `ila_host_alloc` (freed by a finalizer): the library copies such arrays by DMA
with no staging.  Ordinary `Vector`s work too — they go through the library's pinned staging ring — so this only matters
for result arrays that are produced over and over (the batch wrappers below allocate their outputs this way).
"""
function PinnedVector(::Type{T}, n::Integer) where {T}
    p = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ila_host_alloc, lib), Cint, (Ptr{Ptr{Cvoid}}, Int64), p, max(n, 1) * sizeof(T)))
    v = unsafe_wrap(Array, Ptr{T}(p[]), n; own=false)
    finalizer(_ -> ccall((:ila_host_free, lib), Cint, (Ptr{Cvoid},), p[]), v)
    v
end
"Pin an existing array in place for the lifetime of `f()` (cudaHostRegister costs ~10 ms/MB: only for buffers reused many times)."
function with_pinned(f, a::Array)
    check(ccall((:ila_host_register, lib), Cint, (Ptr{Cvoid}, Int64), a, sizeof(a)))
    try
        GC.@preserve a f()
    finally
        ccall((:ila_host_unregister, lib), Cint, (Ptr{Cvoid},), a)
    end
end

"Rebuild the reference's exception from a per-instance status (SURVEY §5)."
function throw_status(st::Integer, what)
    st == 0 && return nothing
    st == 1 && throw(DomainError(what, STATUS_MSG[1]))
    st == 2 && throw(DomainError(what, STATUS_MSG[2]))
    st == 3 && error("Did not converge")                                   # src/infql.jl:181 / capacity reached
    st == 4 && throw(PosDefException(1))                                   # pbtrf info > 0 (discarded upstream, infcholesky.jl:49)
    st == 5 && error("Not-a-number encounted")                             # src/infqr.jl:324
    error("libila_b200: status $st")
end

const STATUS_MSG = Dict(
    1 => "QL factorization does not exist. This could indicate that the operator is not Fredholm or that the dimension of the kernel exceeds that of the co-kernel. Try again with the adjoint.",
    2 => "Real-valued QL factorization does not exist. Try ql(complex(A)) to see if a complex-valued QL factorization exists.")

"""
    ql_L11(datacolumn, λs; branch_rank=1, ngpus=1) -> (L11, status)

`datacolumn` is `A.data.args[1]` of an `InfToeplitz` with bandwidths (l, 1): `[a₊₁, a₀, a₋₁, …, a₋ₗ]`
(src/banded/infqltoeplitz.jl:59 reverses it; the library does that internally).
Replaces `tail_de` + `ql_X!` + `ql_hessenberg` (src/banded/infqltoeplitz.jl:7-65) per shift.
"""
function ql_L11(datacolumn::AbstractVector, λs::AbstractVector; branch_rank::Integer=1, ngpus::Integer=1)
    a = Vector{ComplexF64}(datacolumn)
    λ = Vector{ComplexF64}(λs)
    n = length(λ)
    L11 = Vector{ComplexF64}(undef, n)
    status = Vector{Int32}(undef, n)
    l = length(a) - 2
    check(ccall((:ila_ql_toeplitz_l11_c64, lib), Cint,
                (Cint, Cint, Ptr{ComplexF64}, Ptr{ComplexF64}, Int64, Cint, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{Int32}, Cint),
                l, 1, a, λ, n, branch_rank, L11, C_NULL, status, ngpus))
    L11, status
end

"""
    ql_L11_real(datacolumn, λs; branch_rank=1, ngpus=1) -> (L11::Vector{Float64}, status::Vector{UInt8})

Compact form (`ila_ql_toeplitz_l11re_c64`): `L[1,1]` of a ComplexF64 factorization is real (`ql_X!`'s k = 1 reflector
leaves `-ν`), so 8 B + 1 B per shift come back instead of 16 + 4.  Entries with `status != 0` are NaN with the status
in the NaN payload.  Outputs are pinned arrays.
"""
function ql_L11_real(datacolumn::AbstractVector, λs::AbstractVector; branch_rank::Integer=1, ngpus::Integer=1)
    a = Vector{ComplexF64}(datacolumn)
    λ = λs isa Vector{ComplexF64} ? λs : Vector{ComplexF64}(λs)
    n = length(λ)
    L11 = PinnedVector(Float64, n); status = PinnedVector(UInt8, n)
    check(ccall((:ila_ql_toeplitz_l11re_c64, lib), Cint,
                (Cint, Cint, Ptr{ComplexF64}, Ptr{ComplexF64}, Int64, Cint, Ptr{Float64}, Ptr{UInt8}, Cint),
                length(a) - 2, 1, a, λ, n, branch_rank, L11, status, ngpus))
    L11, status
end

"""
    ql_tail(datacolumn, λs; branch_rank=1, ngpus=1) -> (X::Array{T,3} (2 × m × n), τ::Matrix{T} (2 × n), status)

The whole fixed-point record of `ql_hessenberg(::InfToeplitz)` (src/banded/infqltoeplitz.jl:56-65) per shift: `X` after
`ql_X!` and `τ`, from which first L column = `[X[1,m-1]; X[2,m-1:-1:1]]`, tail L column = `X[2,m:-1:1]` and the 2×2 block
`QL(X[:,m-1:m], τ).Q` follow (`:63-64`).  `T` is `Float64` when symbol and shifts are real (the `T<:Real` method, `:7-16`,
with the sign rule documented in include/ila_b200.h), else `ComplexF64`.
"""
function ql_tail(datacolumn::AbstractVector, λs::AbstractVector; branch_rank::Integer=1, ngpus::Integer=1)
    T = (eltype(datacolumn) <: Real && eltype(λs) <: Real) ? Float64 : ComplexF64
    a = Vector{T}(datacolumn); λ = Vector{T}(λs); n = length(λ); m = length(a)
    L11 = Vector{T}(undef, n); tail = Matrix{T}(undef, 2m + 2, n); status = Vector{Int32}(undef, n)
    sym = T === Float64 ? :ila_ql_toeplitz_l11_f64 : :ila_ql_toeplitz_l11_c64
    check(ccall((sym, lib), Cint, (Cint, Cint, Ptr{T}, Ptr{T}, Int64, Cint, Ptr{T}, Ptr{T}, Ptr{Int32}, Cint),
                m - 2, 1, a, λ, n, branch_rank, L11, tail, status, ngpus))
    X = Array{T,3}(undef, 2, m, n)
    for i in 1:n, j in 1:m                       # record = X[1,1:m], X[2,1:m], τ[1], τ[2]
        X[1, j, i] = tail[j, i]; X[2, j, i] = tail[m + j, i]
    end
    X, tail[2m+1:2m+2, :], status
end

"Scalar drop-in: rethrows the reference's DomainError (src/banded/infqltoeplitz.jl:12-13,23)."
function ql_L11(datacolumn::AbstractVector, λ::Number; kw...)
    L, st = ql_L11(datacolumn, [λ]; kw...)
    st[1] == 0 || throw(DomainError(λ, get(STATUS_MSG, Int(st[1]), "status $(st[1])")))
    L[1]
end

"""
    qr_solve(l, u, p, q, r, b; shift=nothing, head=nothing, tol=floatmin(Float64), colgrowth=1000, ncap, ngpus=1)

Batched `A_i \\ b_i` for ∞ banded operators with affine-in-k diagonals (include/ila_b200.h).  `p,q,r` are
`batch × (l+u+1)` (row-major on the C side: pass `permutedims`), `b` is `batch × nb`.
Replaces `qr(A)` → `adaptiveqr` (src/infqr.jl:163-174) and `ldiv!(F.R, F.Q'*b)` (src/infqr.jl:236-274,350-355).
Returns `(x::Vector{Vector{Float64}}, ncols, nconv, status)`.
"""
function qr_solve(l::Integer, u::Integer, p::Matrix{Float64}, q::Matrix{Float64}, r::Union{Nothing,Matrix{Float64}},
                  b::Matrix{Float64}; shift=nothing, head=nothing, tol=floatmin(Float64), colgrowth=1000,
                  ncap::Integer, ngpus::Integer=1)
    batch = size(p, 1)
    pT, qT, bT = permutedims(p), permutedims(q), permutedims(b)          # instance-major for C
    rT = r === nothing ? C_NULL : permutedims(r)
    hp = head === nothing ? 0 : size(head, 2)
    hT = head === nothing ? C_NULL : permutedims(head)
    sh = shift === nothing ? C_NULL : Vector{Float64}(shift)
    xcap = batch * ncap
    x = Vector{Float64}(undef, xcap)
    xoff = Vector{Int64}(undef, batch + 1); ncols = Vector{Int64}(undef, batch); nconv = similar(ncols)
    status = Vector{Int32}(undef, batch)
    check(ccall((:ila_qr_banded_solve_f64, lib), Cint,
                (Cint, Cint, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint, Ptr{Float64}, Cint,
                 Ptr{Float64}, Float64, Cint, Int64, Ptr{Int64}, Ptr{Float64}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int64},
                 Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Cint),
                l, u, batch, pT, qT, rT, sh, hp, hT, size(b, 2), bT, tol, colgrowth, ncap, C_NULL, x, xcap, xoff, ncols,
                nconv, C_NULL, C_NULL, C_NULL, status, ngpus))
    [x[xoff[i]+1:xoff[i+1]] for i in 1:batch], ncols, nconv, status
end

"""
    absl11_grid(datacolumn, x, y; branch_rank=1, ngpus=1) -> Matrix{Float64}

The example's `ℓ11.(Ref(A), x' .+ y.*im)` (examples/toeplitz.jl:14-33): `abs(L[1,1])`, or `-1.0` where the QL does not
exist.  Shifts are formed on the device; the result is `length(y) × length(x)` (row-major on the C side = the
transpose of what Julia stores, hence the final `permutedims`).
"""
function absl11_grid(datacolumn::AbstractVector, x::AbstractVector{<:Real}, y::AbstractVector{<:Real};
                     branch_rank::Integer=1, ngpus::Integer=1)
    a = Vector{ComplexF64}(datacolumn); xs = Vector{Float64}(x); ys = Vector{Float64}(y)
    out = Matrix{Float64}(undef, length(xs), length(ys))
    check(ccall((:ila_ql_toeplitz_absl11_grid_c64, lib), Cint,
                (Cint, Cint, Ptr{ComplexF64}, Ptr{Float64}, Cint, Ptr{Float64}, Int64, Cint, Ptr{Float64}, Cint),
                length(a) - 2, 1, a, xs, length(xs), ys, length(ys), branch_rank, out, ngpus))
    permutedims(out)
end

"""
    ql_solve(datacolumn, λs, b, nout; branch_rank=1, ngpus=1) -> (X::Matrix, status)

`X[i, 1:nout] = (ql(A - λs[i]*I) \\ b)[1:nout]` for a pure Toeplitz `A` (u == 1) and a finitely supported `b`
(`length(b) ≤ 16`).  Replaces `ldiv!(F.L, lmul!(F.Q', b))` (src/infql.jl:266-289, src/banded/hessenbergq.jl:125-135).
The library writes `X` column-major n × nout — exactly Julia's `Matrix` layout.
"""
function ql_solve(datacolumn::AbstractVector, λs::AbstractVector, b::AbstractVector, nout::Integer;
                  branch_rank::Integer=1, ngpus::Integer=1)
    a = Vector{ComplexF64}(datacolumn); λ = Vector{ComplexF64}(λs); bb = Vector{ComplexF64}(b)
    n = length(λ)
    X = Matrix{ComplexF64}(undef, n, nout); status = Vector{Int32}(undef, n)
    check(ccall((:ila_ql_toeplitz_solve_c64, lib), Cint,
                (Cint, Cint, Ptr{ComplexF64}, Ptr{ComplexF64}, Int64, Cint, Cint, Ptr{ComplexF64}, Int64, Ptr{ComplexF64},
                 Ptr{Int32}, Cint),
                length(a) - 2, 1, a, λ, n, branch_rank, length(bb), bb, nout, X, status, ngpus))
    X, status
end

"""
    ql_pert_L11(head, datacolumn, λs; branch_rank=1, ngpus=1) -> (L11, status)

`ql(A - λ*I).L[1,1]` for a perturbed Toeplitz operator: `head` is the p × p top-left corner of `A` (unshifted), the
tail symbol as in `ql_L11`.  Replaces `ql_hessenberg!` (src/infql.jl:70-93).
"""
function ql_pert_L11(head::AbstractMatrix, datacolumn::AbstractVector, λs::AbstractVector; branch_rank::Integer=1,
                     ngpus::Integer=1)
    a = Vector{ComplexF64}(datacolumn); λ = Vector{ComplexF64}(λs); h = Matrix{ComplexF64}(head)
    n = length(λ); p = size(h, 1)
    L11 = Vector{ComplexF64}(undef, n); status = Vector{Int32}(undef, n)
    check(ccall((:ila_ql_perttoeplitz_l11_c64, lib), Cint,
                (Cint, Cint, Cint, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{ComplexF64}, Int64, Cint, Ptr{ComplexF64},
                 Ptr{ComplexF64}, Ptr{Int32}, Cint),
                length(a) - 2, 1, p, h, a, λ, n, branch_rank, L11, C_NULL, status, ngpus))
    L11, status
end

"""
    blocktri_L11(dl_head, d_head, du_head, c, a, b, energies; atol=1e-10, maxit=1_000_000, ngpus=1)

`ql(A - x*I).L[1,1]` over `energies` for `A = BlockTridiagonal(Vcat(dl_head, Fill(c,∞)), Vcat(d_head, Fill(a,∞)),
Vcat(du_head, Fill(b,∞)))` with n × n real blocks (n ≤ 4).  Replaces `_blocktripert_ql` / `blocktailiterate`
(src/infql.jl:166-209).  Returns `(L11, iters, status)`; status 3 = "Did not converge" (src/infql.jl:181).
(`ila_ql_blocktri_l11_c64` is the same call with `ComplexF64` blocks / energies.)
"""
function blocktri_L11(dl_head::Vector{Matrix{Float64}}, d_head::Vector{Matrix{Float64}}, du_head::Vector{Matrix{Float64}},
                      c::Matrix{Float64}, a::Matrix{Float64}, b::Matrix{Float64}, energies::AbstractVector{<:Real};
                      atol::Real=1e-10, maxit::Integer=1_000_000, ngpus::Integer=1)
    n = size(c, 1); e = Vector{Float64}(energies); m = length(e)
    cat3(v) = isempty(v) ? Float64[] : reduce(vcat, vec.(v))      # blocks are column-major already
    L11 = Vector{Float64}(undef, m); iters = Vector{Int32}(undef, m); status = Vector{Int32}(undef, m)
    check(ccall((:ila_ql_blocktri_l11_f64, lib), Cint,
                (Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint, Ptr{Float64}, Cint, Ptr{Float64}, Cint, Ptr{Float64},
                 Ptr{Float64}, Int64, Float64, Cint, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Cint),
                n, c, a, b, length(dl_head), cat3(dl_head), length(d_head), cat3(d_head), length(du_head), cat3(du_head),
                e, m, atol, maxit, L11, iters, status, ngpus))
    L11, iters, status
end

"Bessel recurrence of README.md:41-43: `A \\ [besselj(1,z); Zeros(∞)]` for many z at once."
function bessel_solve(zs::AbstractVector{<:Real}, b1::AbstractVector{<:Real}; ngpus::Integer=1)
    batch = length(zs)
    p = repeat([1.0 0.0 1.0], batch); q = repeat([0.0 -2.0 0.0], batch); r = ones(batch, 3); r[:, 2] .= zs
    ncap = ceil(Int, maximum(zs) + 90cbrt(maximum(zs)) + 2100)
    qr_solve(1, 1, p, q, r, reshape(Vector{Float64}(b1), batch, 1); ncap=ncap, ngpus=ngpus)
end

"""
    triconj(U, X, n; ngpus=1) -> Y (3 x n: dl, d, du)

`Tridiagonal(U*X/U)` for one operator: `U` is `nbands x m` (rows `U[k,k]`, `U[k,k+1]`[, `U[k,k+2]`]), `X` is `3 x m`
(`dl`, `d`, `du`), `m >= n+1`.  Replaces `upper_mul_tri_triview!` + `tri_mul_invupper_triview!` as driven by
`resizedata!(::TridiagonalConjugationData, n)` (src/banded/tridiagonalconjugation.jl:10-215, 260-282).
"""
function triconj(U::Matrix{Float64}, X::Matrix{Float64}, n::Integer; ngpus::Integer=1)
    Ut = permutedims(U); Xt = permutedims(X)          # column-major m x nbands == the ABI's band-major rows
    m = size(Ut, 1)
    Y = Matrix{Float64}(undef, n, 3)
    rc = ccall((:ila_triconj_bands_f64, lib), Cint,
               (Cint, Int64, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Cint, Ptr{Float64}, Ptr{Float64}, Cint),
               size(U, 1), 1, n, Ut, m, Xt, size(Xt, 1), 1, Y, C_NULL, ngpus)
    check(rc)
    permutedims(Y)
end

"""
    blockbanded_solve(T1, T2, α, β, λs; b=[1.0], tol=1e-30, colgrowth=300, maxblocks=64, ngpus=1) -> (x, ncols, status)

`(α KronTrav(T1, Eye) + β KronTrav(Eye, T2) - λ I) \ [b; zeros(∞)]` for every λ in `λs` by the adaptive block-banded QR
(src/infqr.jl:23-28, 68-86, 303-337, 358-366).  `T1`, `T2`: `3 x m` bands (`dl`, `d`, `du`), `m >= maxblocks+1`.
Column `i` of `x` is the zero-padded solution for `λs[i]`, `ncols[i]` its length.
"""
function blockbanded_solve(T1::Matrix{Float64}, T2::Matrix{Float64}, α::Real, β::Real, λs::AbstractVector{<:Real};
                           b::Vector{Float64}=[1.0], tol::Real=1e-30, colgrowth::Integer=300, maxblocks::Integer=64,
                           ngpus::Integer=1)
    batch = length(λs); xld = maxblocks * (maxblocks + 1) ÷ 2
    T1t = permutedims(T1); T2t = permutedims(T2)
    x = Matrix{Float64}(undef, xld, batch); ncols = Vector{Int64}(undef, batch)
    nblocks = Vector{Int32}(undef, batch); status = Vector{Int32}(undef, batch)
    rc = ccall((:ila_qr_blockbanded_solve_f64, lib), Cint,
               (Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Cint, Ptr{Float64}, Float64,
                Cint, Cint, Ptr{Float64}, Int64, Ptr{Int64}, Ptr{Int32}, Ptr{Float64}, Ptr{Int32}, Cint),
               batch, fill(Float64(α), batch), fill(Float64(β), batch), Vector{Float64}(λs), T1t, T2t, size(T1t, 1), length(b),
               repeat(b, batch), tol, colgrowth, maxblocks, x, xld, ncols, nblocks, C_NULL, status, ngpus)
    check(rc)
    x, ncols, status
end

"""
    almostbanded_solve(lD, uD, dg, bg, b, shifts; tol=floatmin(Float64), colgrowth=1000, ncap=6000, ngpus=1) -> (x, ncols, status)

`Vcat(B, D - λ I) \ [b; zeros(∞)]` for every λ in `shifts` (src/infqr.jl:15-21, 50-66, 236-274, 350-355): `dg` is the
`(lD+uD+1) x 3` matrix of polynomial generators of D's diagonals (+uD first, `g0 + g1 t + g2 t^2` in the row index of D), `bg`
the `r x 4` matrix of boundary rows `s^j (a + b j + c j^2)`.  Column `i` of `x` is the zero-padded solution for `shifts[i]`.
"""
function almostbanded_solve(lD::Integer, uD::Integer, dg::Matrix{Float64}, bg::Matrix{Float64}, b::Vector{Float64},
                            shifts::AbstractVector{<:Real}; tol::Real=floatmin(Float64), colgrowth::Integer=1000,
                            ncap::Integer=6000, ngpus::Integer=1)
    batch = length(shifts); r = size(bg, 1)
    dgt = repeat(vec(permutedims(dg)), batch); bgt = repeat(vec(permutedims(bg)), batch)     # operator-major, row-major blocks
    x = Matrix{Float64}(undef, ncap, batch); ncols = Vector{Int64}(undef, batch); nconv = Vector{Int64}(undef, batch)
    status = Vector{Int32}(undef, batch)
    rc = ccall((:ila_qr_almostbanded_solve_f64, lib), Cint,
               (Cint, Cint, Cint, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint, Ptr{Float64}, Float64, Cint, Int64,
                Ptr{Float64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int32}, Cint),
               lD, uD, r, batch, dgt, bgt, Vector{Float64}(shifts), length(b), repeat(b, batch), tol, colgrowth, ncap, x, ncols,
               nconv, status, ngpus)
    check(rc)
    x, ncols, status
end


# ---- the operator description shared by the adaptive QR / Cholesky entries -------------------------------------------
# instance-major (C row-major) flattening of a batch × k Julia matrix
_rm(M::AbstractMatrix) = vec(permutedims(Matrix{Float64}(M)))
_rm(::Nothing) = C_NULL

"""
    chol_solve(u, p, q, r, b; shift, head, tol, colgrowth, ncap, ngpus) -> (x::Vector{Vector{Float64}}, ncols, nconv, status)

Batched `cholesky(S_i) \ b_i` for ∞ symmetric positive definite banded operators (upper bandwidth `u`, generators of the
UPPER bands only, data-row order +u … 0).  Replaces `adaptivecholesky` / `partialcholesky!` / the adaptive `L \ b` and
`U \ b` (src/infcholesky.jl:42-59, 71-75, 94-140).  status 4 = not positive definite.
"""
function chol_solve(u::Integer, p::Matrix{Float64}, q::Matrix{Float64}, r::Union{Nothing,Matrix{Float64}}, b::Matrix{Float64};
                    shift=nothing, head=nothing, tol=floatmin(Float64), colgrowth=1000, ncap::Integer, ngpus::Integer=1)
    batch = size(p, 1); xcap = batch * ncap
    x = Vector{Float64}(undef, xcap); xoff = Vector{Int64}(undef, batch + 1)
    ncols = Vector{Int64}(undef, batch); nconv = similar(ncols); status = Vector{Int32}(undef, batch)
    hp = head === nothing ? 0 : size(head, 2)
    sh = shift === nothing ? C_NULL : Vector{Float64}(shift)
    check(ccall((:ila_chol_banded_solve_f64, lib), Cint,
                (Cint, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint, Ptr{Float64}, Cint, Ptr{Float64},
                 Float64, Cint, Int64, Ptr{Int64}, Ptr{Float64}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64},
                 Ptr{Float64}, Ptr{Int32}, Cint),
                u, batch, _rm(p), _rm(q), _rm(r), sh, hp, head === nothing ? C_NULL : _rm(head), size(b, 2), _rm(b), tol,
                colgrowth, ncap, C_NULL, x, xcap, xoff, ncols, nconv, C_NULL, C_NULL, status, ngpus))
    [x[xoff[i]+1:xoff[i+1]] for i in 1:batch], ncols, nconv, status
end

"""
    qr_solve_c64(l, u, p, q, r, b; shift, head, tol, colgrowth, ncap, ngpus)

Complex eltype of `qr_solve` (`ila_qr_banded_solve_c64`): `(A - λ_i I) \ b` by adaptive QR with complex shifts / operators.
"""
function qr_solve_c64(l::Integer, u::Integer, p::Matrix{ComplexF64}, q::Matrix{ComplexF64}, r::Union{Nothing,Matrix{Float64}},
                      b::Matrix{ComplexF64}; shift=nothing, head=nothing, tol=floatmin(Float64), colgrowth=1000,
                      ncap::Integer, ngpus::Integer=1)
    batch = size(p, 1); xcap = batch * ncap
    cm(M) = vec(permutedims(M))
    x = Vector{ComplexF64}(undef, xcap); xoff = Vector{Int64}(undef, batch + 1)
    ncols = Vector{Int64}(undef, batch); nconv = similar(ncols); status = Vector{Int32}(undef, batch)
    hp = head === nothing ? 0 : size(head, 2)
    check(ccall((:ila_qr_banded_solve_c64, lib), Cint,
                (Cint, Cint, Int64, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{Float64}, Ptr{ComplexF64}, Cint, Ptr{ComplexF64}, Cint,
                 Ptr{ComplexF64}, Float64, Cint, Int64, Ptr{Int64}, Ptr{ComplexF64}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int64},
                 Ptr{Int32}, Cint),
                l, u, batch, cm(p), cm(q), r === nothing ? C_NULL : _rm(r), shift === nothing ? C_NULL : Vector{ComplexF64}(shift),
                hp, head === nothing ? C_NULL : cm(Matrix{ComplexF64}(head)), size(b, 2), cm(b), tol, colgrowth, ncap, C_NULL, x,
                xcap, xoff, ncols, nconv, status, ngpus))
    [x[xoff[i]+1:xoff[i+1]] for i in 1:batch], ncols, nconv, status
end

"""
    ql_pert_solve(head, datacolumn, λs, b, nout; branch_rank=1, ngpus=1) -> (X::Matrix{ComplexF64} (n × nout), status)

`(ql(A - λ I) \ b)[1:nout]` for a perturbed Toeplitz `A` (`ila_ql_perttoeplitz_solve_c64`): `ql_hessenberg!` assembly
(src/infql.jl:70-93) + `ldiv!(F.L, lmul!(F.Q', b))` (src/infql.jl:266-289).  Pinned by test/test_infql.jl:147-160.
"""
function ql_pert_solve(head::AbstractMatrix, datacolumn::AbstractVector, λs::AbstractVector, b::AbstractVector, nout::Integer;
                       branch_rank::Integer=1, ngpus::Integer=1)
    a = Vector{ComplexF64}(datacolumn); λ = Vector{ComplexF64}(λs); h = Matrix{ComplexF64}(head); bb = Vector{ComplexF64}(b)
    n = length(λ); X = Matrix{ComplexF64}(undef, n, nout); status = Vector{Int32}(undef, n)
    check(ccall((:ila_ql_perttoeplitz_solve_c64, lib), Cint,
                (Cint, Cint, Cint, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{ComplexF64}, Int64, Cint, Cint, Ptr{ComplexF64}, Int64,
                 Ptr{ComplexF64}, Ptr{Int32}, Cint),
                length(a) - 2, 1, size(h, 1), h, a, λ, n, branch_rank, length(bb), bb, nout, X, status, ngpus))
    X, status
end

"""
    revchol_tridiag(ga, gb, n; shift, a_head, b_head, ngpus) -> (la, lb, N, status)

Batched `reversecholesky(SymTridiagonal(a, b))` leading `n` entries (`ila_revchol_tridiag_f64`;
src/banded/infreversecholeskytridiagonal.jl:66-141).  `ga`, `gb`: batch × 5 rational generators `(c0, p, q, r, s)` of the
diagonal / off-diagonal, value(k) = c0 + (p + q k)/(r + s k).  `la`, `lb` are n × batch.
"""
function revchol_tridiag(ga::Matrix{Float64}, gb::Matrix{Float64}, n::Integer; shift=nothing, a_head=nothing, b_head=nothing,
                         ngpus::Integer=1)
    batch = size(ga, 1)
    la = Matrix{Float64}(undef, n, batch); lb = similar(la); N = Vector{Int64}(undef, batch); status = Vector{Int32}(undef, batch)
    hp = a_head === nothing ? 0 : length(a_head)
    check(ccall((:ila_revchol_tridiag_f64, lib), Cint,
                (Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64},
                 Ptr{Int64}, Ptr{Int32}, Cint),
                batch, _rm(ga), _rm(gb), shift === nothing ? C_NULL : Vector{Float64}(shift), hp,
                a_head === nothing ? C_NULL : Vector{Float64}(a_head), b_head === nothing ? C_NULL : Vector{Float64}(b_head), n, la,
                lb, N, status, ngpus))
    la, lb, N, status
end

"""
    finite_section_ql(l, u, g, shifts; j=50, tol=eps(), maxN=10000, head=nothing, ngpus=1) -> (Lband, V, τ, N, status)

Batched `ql(A - λ I, tol)` for ∞ banded operators WITHOUT a Toeplitz tail (`ila_ql_finite_section_f64/_c64`;
src/infql.jl:385-462).  `g`: (l+u+1) × 5 rational generators per diagonal (+u first), shared by the batch; `shifts` real or
complex.  `Lband[c+1, k, i] = L[k, k-c]`, `V[r, k, i]`, `τ[k, i]` for the leading `j` columns.
"""
function finite_section_ql(l::Integer, u::Integer, g::Matrix, shifts::AbstractVector; j::Integer=50, tol::Real=eps(Float64),
                           maxN::Integer=10000, head=nothing, ngpus::Integer=1)
    T = (eltype(g) <: Real && eltype(shifts) <: Real && (head === nothing || eltype(head) <: Real)) ? Float64 : ComplexF64
    batch = length(shifts); nd = l + u + 1
    gg = repeat(vec(permutedims(Matrix{T}(g))), batch)
    Lb = Array{T,3}(undef, nd, j, batch); V = Array{T,3}(undef, u, j, batch); τ = Matrix{T}(undef, j, batch)
    N = Vector{Int64}(undef, batch); status = Vector{Int32}(undef, batch)
    hp = head === nothing ? 0 : size(head, 2)
    sym = T === Float64 ? :ila_ql_finite_section_f64 : :ila_ql_finite_section_c64
    check(ccall((sym, lib), Cint,
                (Cint, Cint, Int64, Ptr{T}, Ptr{T}, Cint, Ptr{T}, Int64, Float64, Int64, Ptr{T}, Ptr{T}, Ptr{T}, Ptr{Int64}, Ptr{Int32}, Cint),
                l, u, batch, gg, Vector{T}(shifts), hp, head === nothing ? C_NULL : vec(permutedims(Matrix{T}(head))), j, tol, maxN,
                Lb, V, τ, N, status, ngpus))
    Lb, V, τ, N, status
end

"""
    biconj(U, C, n; upper=true, ngpus=1) -> (dv, ev)

`BidiagonalConjugation(U, X, V, uplo)` bands (`ila_biconj_bands_f64`; src/banded/bidiagonalconjugation.jl:19-60) from the
tridiagonal parts of `U` and `C = X*V`, each 3 × m (`[k+2,k+1]`, diagonal, `[k+1,k+2]`), m ≥ n+1.
"""
function biconj(U::Matrix{Float64}, Cm::Matrix{Float64}, n::Integer; upper::Bool=true, ngpus::Integer=1)
    Ut = permutedims(U); Ct = permutedims(Cm); ld = size(Ut, 1)
    dv = Vector{Float64}(undef, n); ev = Vector{Float64}(undef, n)
    check(ccall((:ila_biconj_bands_f64, lib), Cint,
                (Cint, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Cint),
                upper ? 1 : 0, 1, n, Ut, Ct, ld, dv, ev, ngpus))
    dv, ev
end

"""
    chol_triconj(u, p, q, X, n; shift, head, ngpus) -> (Y::Array{Float64,3} (n × 3 × batch), status)

`TridiagonalConjugation(cholesky(S_i).U, X)` in one call, `U` never leaves the device (`ila_chol_triconj_f64`;
src/banded/tridiagonalconjugation.jl:224-282 behind src/infcholesky.jl:42-59).  `X`: 3 × m bands shared by the batch.
"""
function chol_triconj(u::Integer, p::Matrix{Float64}, q::Matrix{Float64}, X::Matrix{Float64}, n::Integer; shift=nothing,
                      head=nothing, ngpus::Integer=1)
    batch = size(p, 1); Xt = permutedims(X); hp = head === nothing ? 0 : size(head, 2)
    Y = Array{Float64,3}(undef, n, 3, batch); status = Vector{Int32}(undef, batch)
    check(ccall((:ila_chol_triconj_f64, lib), Cint,
                (Cint, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint, Ptr{Float64}, Int64, Ptr{Float64}, Int64,
                 Cint, Ptr{Float64}, Ptr{Int32}, Cint),
                u, batch, _rm(p), _rm(q), C_NULL, shift === nothing ? C_NULL : Vector{Float64}(shift), hp,
                head === nothing ? C_NULL : _rm(head), n, Xt, size(Xt, 1), 1, Y, status, ngpus))
    Y, status
end

"Complex eltype of `blocktri_L11` (`ila_ql_blocktri_l11_c64`, e.g. `A - 5im*I`, test/test_periodic.jl:20-26)."
function blocktri_L11(dl_head::Vector{Matrix{ComplexF64}}, d_head::Vector{Matrix{ComplexF64}}, du_head::Vector{Matrix{ComplexF64}},
                      c::Matrix{ComplexF64}, a::Matrix{ComplexF64}, b::Matrix{ComplexF64}, energies::AbstractVector;
                      atol::Real=1e-10, maxit::Integer=1_000_000, ngpus::Integer=1)
    n = size(c, 1); e = Vector{ComplexF64}(energies); m = length(e)
    cat3(v) = isempty(v) ? ComplexF64[] : reduce(vcat, vec.(v))
    L11 = Vector{ComplexF64}(undef, m); iters = Vector{Int32}(undef, m); status = Vector{Int32}(undef, m)
    T = ComplexF64
    check(ccall((:ila_ql_blocktri_l11_c64, lib), Cint,
                (Cint, Ptr{T}, Ptr{T}, Ptr{T}, Cint, Ptr{T}, Cint, Ptr{T}, Cint, Ptr{T}, Ptr{T}, Int64, Float64, Cint, Ptr{T},
                 Ptr{Int32}, Ptr{Int32}, Cint),
                n, c, a, b, length(dl_head), cat3(dl_head), length(d_head), cat3(d_head), length(du_head), cat3(du_head), e, m,
                atol, maxit, L11, iters, status, ngpus))
    L11, iters, status
end

"""
    apply_Q(datacolumn, λs, b, nout; adjoint=false, branch_rank=1, ngpus=1) -> (Y::Matrix (n × nout), nstop, status)

`Q*b` (`adjoint=false`: `LowerHessenbergQ` apply with the reference's early exit, src/banded/hessenbergq.jl:111-123; `nstop[i]` =
the n at which it fired, 0 if beyond `nout`) or `Q'*b` (`:125-135`) for the ∞ Hessenberg Q of `ql(A - λs[i]*I)`.
"""
function apply_Q(datacolumn::AbstractVector, λs::AbstractVector, b::AbstractVector, nout::Integer; adjoint::Bool=false,
                 branch_rank::Integer=1, ngpus::Integer=1)
    a = Vector{ComplexF64}(datacolumn); λ = Vector{ComplexF64}(λs); bb = Vector{ComplexF64}(b); n = length(λ)
    Y = Matrix{ComplexF64}(undef, n, nout); nstop = Vector{Int64}(undef, n); status = Vector{Int32}(undef, n)
    check(ccall((:ila_ql_toeplitz_applyq_c64, lib), Cint,
                (Cint, Cint, Ptr{ComplexF64}, Ptr{ComplexF64}, Int64, Cint, Cint, Cint, Ptr{ComplexF64}, Int64, Ptr{ComplexF64},
                 Ptr{Int64}, Ptr{Int32}, Cint),
                length(a) - 2, 1, a, λ, n, branch_rank, adjoint ? 1 : 0, length(bb), bb, nout, Y, adjoint ? C_NULL : nstop, status, ngpus))
    Y, nstop, status
end

"""
    blocktri_factors(dl_head, d_head, du_head, c, a, b, energies; rhs=nothing, nout=0, atol=1e-10, maxit=1_000_000, ngpus=1)

The whole block QL per energy (`ila_ql_blocktri_factors_f64`): `(L11, P (2n × 3n × m), τtail (n × m), head (Nn × Nn × m),
headτ (Nn × m), qtb (nout × m) or nothing, iters, status)` — the records `QL(_BlockSkylineMatrix(Vcat(BB.data, Fill(tail
column))), Vcat(F.τ, Fill(τ,∞)))` is assembled from (src/infql.jl:199-204), and `(Q'rhs)[1:nout]` (`:213-243`).
"""
function blocktri_factors(dl_head::Vector{Matrix{Float64}}, d_head::Vector{Matrix{Float64}}, du_head::Vector{Matrix{Float64}},
                          c::Matrix{Float64}, a::Matrix{Float64}, b::Matrix{Float64}, energies::AbstractVector{<:Real};
                          rhs=nothing, nout::Integer=0, atol::Real=1e-10, maxit::Integer=1_000_000, ngpus::Integer=1)
    n = size(c, 1); e = Vector{Float64}(energies); m = length(e)
    N = max(length(du_head) + 1, length(d_head), length(dl_head)); hn = N * n
    cat3(v) = isempty(v) ? Float64[] : reduce(vcat, vec.(v))
    L11 = Vector{Float64}(undef, m); P = Array{Float64,3}(undef, 2n, 3n, m); τt = Matrix{Float64}(undef, n, m)
    H = Array{Float64,3}(undef, hn, hn, m); Hτ = Matrix{Float64}(undef, hn, m)
    iters = Vector{Int32}(undef, m); status = Vector{Int32}(undef, m)
    rb = rhs === nothing ? Float64[] : Vector{Float64}(rhs)
    qtb = rhs === nothing ? nothing : Matrix{Float64}(undef, nout, m)
    check(ccall((:ila_ql_blocktri_factors_f64, lib), Cint,
                (Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint, Ptr{Float64}, Cint, Ptr{Float64}, Cint, Ptr{Float64},
                 Ptr{Float64}, Int64, Float64, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint,
                 Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Cint),
                n, c, a, b, length(dl_head), cat3(dl_head), length(d_head), cat3(d_head), length(du_head), cat3(du_head), e, m,
                atol, maxit, L11, P, τt, H, Hτ, length(rb), isempty(rb) ? C_NULL : rb, nout, qtb === nothing ? C_NULL : qtb,
                iters, status, ngpus))
    L11, P, τt, H, Hτ, qtb, iters, status
end

# ---- library plumbing ------------------------------------------------------------------------------------------------
device_count() = Int(ccall((:ila_device_count, lib), Cint, ()))
set_device(d::Integer) = check(ccall((:ila_set_device, lib), Cint, (Cint,), d))
release() = check(ccall((:ila_release, lib), Cint, ()))
"0 auto (pageable arrays are staged through the library's pinned ring), 1 plain cudaMemcpyAsync, 2 always staged"
set_host_copy_mode(m::Integer) = check(ccall((:ila_set_host_copy_mode, lib), Cint, (Cint,), m))

end # module
