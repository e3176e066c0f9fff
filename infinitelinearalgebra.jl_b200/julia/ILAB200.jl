# ILAB200.jl — ccall shim over libila_b200.so (include/ila_b200.h).
#
# NOT EXECUTED in this repository's CI: Julia is not installed in the build image nor on the GPU box (SURVEY §0.2).
# It documents the binding a maintainer adds to InfiniteLinearAlgebra.jl so that the existing API stays a drop-in:
#   ql(A - λ*I).L[1,1]            -> ILAB200.ql_L11(A, [λ])[1]
#   ℓ11.(Ref(A), x' .+ y.*im)     -> ILAB200.ql_L11(A, vec(x' .+ y.*im))        (examples/toeplitz.jl:14-33)
#   A \ b  (∞ banded, affine bands) -> ILAB200.qr_solve(...)                     (src/infqr.jl:370-374)
# Exceptions of the reference are rebuilt from status[] (SURVEY §5).
module ILAB200

using LinearAlgebra

const lib = get(ENV, "ILA_B200_LIB", "libila_b200.so")

struct IlaError <: Exception
    code::Cint
    msg::String
end
lasterror() = unsafe_string(ccall((:ila_last_error, lib), Cstring, ()))
check(rc) = rc == 0 || throw(IlaError(rc, lasterror()))

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

"Bessel recurrence of README.md:41-43: `A \\ [besselj(1,z); Zeros(∞)]` for many z at once."
function bessel_solve(zs::AbstractVector{<:Real}, b1::AbstractVector{<:Real}; ngpus::Integer=1)
    batch = length(zs)
    p = repeat([1.0 0.0 1.0], batch); q = repeat([0.0 -2.0 0.0], batch); r = ones(batch, 3); r[:, 2] .= zs
    ncap = ceil(Int, maximum(zs) + 90cbrt(maximum(zs)) + 2100)
    qr_solve(1, 1, p, q, r, reshape(Vector{Float64}(b1), batch, 1); ncap=ncap, ngpus=ngpus)
end

end # module
