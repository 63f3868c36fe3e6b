# ConstrainedControlB200.jl -- Julia shim over libccb.so (include/ccb.h).
#
# Keeps the reference's user-facing surface (`LQR`, `TrackingLQR`, `PID`, `simulate!(mechanism, steps, controller!)`,
# reference src/control/{lqr,lqr_tracking,pid}.jl) for a BATCH of mechanism states and `ccall`s the B200 library for the
# hot path.  NOT TESTED in the build environment (no julia toolchain there); the Python ctypes host
# (constrainedcontrol.jl_b200/control.py) exercises exactly the same entry points.
module ConstrainedControlB200

using ConstrainedDynamics
using ConstrainedDynamics: Mechanism, Body, EqualityConstraint, Storage, Controller, getid, geteqconstraint
using LinearAlgebra

const libccb = get(ENV, "CCB_LIB", joinpath(@__DIR__, "..", "csrc", "libccb.so"))

export BatchLQR, DeviceLQR, DeviceTrackingLQR, BatchPID, BatchStates, replicate!, simulate!

# ------------------------------------------------------------------------------------------------ errors
struct CcbError <: Exception
    code::Int32
    msg::String
end
function check(rc::Integer)
    rc == 0 && return
    throw(CcbError(rc, unsafe_string(ccall((:ccb_last_error, libccb), Cstring, ()))))
end

# ------------------------------------------------------------------------------------------------ mechanism description
# mirrors `struct ccb_mech_desc` (include/ccb.h)
struct MechDesc
    nb::Int32; ne::Int32; nj::Int32; id_scheme::Int32
    dt::Float64; g::Float64; reg::Float64
    mass::Ptr{Float64}; inertia::Ptr{Float64}
    eqc_parent::Ptr{Int32}; eqc_joint_ptr::Ptr{Int32}
    joint_child::Ptr{Int32}; joint_nT::Ptr{Int32}; joint_nR::Ptr{Int32}
    joint_axis::Ptr{Float64}; joint_p1::Ptr{Float64}; joint_p2::Ptr{Float64}; joint_qoff::Ptr{Float64}
end

mutable struct MechHandle
    ptr::Ptr{Cvoid}
    nb::Int; ne::Int; ml::Int
    function MechHandle(mechanism::Mechanism; reg = 0.0)
        bodies = collect(mechanism.bodies); eqcs = collect(mechanism.eqconstraints)
        nb, ne = length(bodies), length(eqcs)
        bindex = Dict(getid(b) => i - 1 for (i, b) in enumerate(bodies))          # position, as `enumerate` in lqr.jl:91
        mass = Float64[b.m for b in bodies]
        inertia = reduce(vcat, [vec(permutedims(Matrix(b.J))) for b in bodies])     # row-major 3x3 per body
        eqc_parent = Int32[get(bindex, e.parentid, -1) for e in eqcs]
        jptr = Int32[0]; jchild = Int32[]; jnT = Int32[]; jnR = Int32[]
        jaxis = Float64[]; jp1 = Float64[]; jp2 = Float64[]; jqoff = Float64[]
        for e in eqcs
            # every ConstrainedDynamics prototype is a (Translational{N1}, Rotational{N2}) pair per child (SURVEY A.3)
            for (k, childid) in enumerate(unique(e.childids))
                tr, rot = joint_pair(e, childid)
                push!(jchild, bindex[childid]); push!(jnT, nrows(tr)); push!(jnR, nrows(rot))
                append!(jaxis, free_axis(tr, rot)); append!(jp1, vertex1(tr)); append!(jp2, vertex2(tr)); append!(jqoff, qoffset(rot))
            end
            push!(jptr, length(jchild))
        end
        desc = Ref(MechDesc(nb, ne, length(jchild), 0, mechanism.Δt, mechanism.g, reg, pointer(mass), pointer(inertia), pointer(eqc_parent),
                            pointer(jptr), pointer(jchild), pointer(jnT), pointer(jnR), pointer(jaxis), pointer(jp1), pointer(jp2), pointer(jqoff)))
        out = Ref{Ptr{Cvoid}}(C_NULL)
        GC.@preserve mass inertia eqc_parent jptr jchild jnT jnR jaxis jp1 jp2 jqoff begin
            check(ccall((:ccb_mech_create, libccb), Cint, (Ref{MechDesc}, Ref{Ptr{Cvoid}}), desc, out))
        end
        dims = zeros(Int32, 8)
        check(ccall((:ccb_mech_dims, libccb), Cint, (Ptr{Cvoid}, Ptr{Int32}), out[], dims))
        h = new(out[], dims[1], dims[2], dims[4])
        finalizer(x -> ccall((:ccb_mech_destroy, libccb), Cint, (Ptr{Cvoid},), x.ptr), h)
        return h
    end
end

# joint introspection helpers: the only ConstrainedDynamics internals the shim needs (field names as of ^0.9.1)
joint_pair(e, childid) = (tr = nothing; rot = nothing;
    for (c, cid) in zip(e.constraints, e.childids)
        cid == childid || continue
        c isa ConstrainedDynamics.Translational && (tr = c)
        c isa ConstrainedDynamics.Rotational && (rot = c)
    end; (tr, rot))
nrows(::Nothing) = Int32(0)
nrows(::ConstrainedDynamics.Joint{T,N}) where {T,N} = Int32(N)
vertex1(tr) = tr === nothing ? zeros(3) : collect(tr.vertices[1])
vertex2(tr) = tr === nothing ? zeros(3) : collect(tr.vertices[2])
qoffset(rot) = rot === nothing ? [1.0, 0, 0, 0] : [rot.qoffset.w, rot.qoffset.x, rot.qoffset.y, rot.qoffset.z]
function free_axis(tr, rot)
    for j in (tr, rot)
        j === nothing && continue
        N = nrows(j)
        N == 2 && return collect(vec(j.V3))        # 1 free direction: the nullspace row
        N == 1 && return collect(vec(j.V3))        # constrained along V3
    end
    return [1.0, 0, 0]
end

# ------------------------------------------------------------------------------------------------ batch of states
"B copies of one mechanism; state[:, i, s] = (xc, qc, vc, ωc) of body i in system s"
mutable struct BatchStates
    mech::MechHandle
    ptr::Ptr{Cvoid}
    B::Int
    function BatchStates(mh::MechHandle, state::Array{Float64,3}; device = 0)
        @assert size(state, 1) == 13 && size(state, 2) == mh.nb
        B = size(state, 3)
        out = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:ccb_batch_create, libccb), Cint, (Ptr{Cvoid}, Int32, Int32, Ref{Ptr{Cvoid}}), mh.ptr, B, device, out))
        b = new(mh, out[], B)
        finalizer(x -> ccall((:ccb_batch_destroy, libccb), Cint, (Ptr{Cvoid},), x.ptr), b)
        check(ccall((:ccb_batch_set_state, libccb), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), b.ptr, state, C_NULL))   # Julia column-major [13][nb][B] == C [B][nb][13]
        return b
    end
end
function getstate(b::BatchStates)
    st = Array{Float64}(undef, 13, b.mech.nb, b.B)
    check(ccall((:ccb_batch_get_state, libccb), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), b.ptr, st, C_NULL))
    return st
end
"state array of a single Mechanism (the reference's `body.state.{xc,qc,vc,ωc}`, examples/lqr_sawyer.jl:16-17)"
function statevector(mechanism::Mechanism)
    st = zeros(13, length(mechanism.bodies))
    for (i, body) in enumerate(mechanism.bodies)
        s = body.state
        st[:, i] = [s.xc; s.qc.w; s.qc.x; s.qc.y; s.qc.z; s.vc; s.ωc]
    end
    return st
end

# ------------------------------------------------------------------------------------------------ controllers
# mirrors `struct ccb_ctrl_desc`
struct CtrlDesc
    kind::Int32; nu::Int32; N::Int32; B::Int32
    eqc::Ptr{Int32}
    K::Ptr{Float64}; K_bstride::Int64
    zd::Ptr{Float64}; zd_bstride::Int64
    Fd::Ptr{Float64}; Fd_bstride::Int64
    U::Ptr{Float64}; U_bstride::Int64
    P::Ptr{Float64}; I::Ptr{Float64}; D::Ptr{Float64}
    goal::Ptr{Float64}; goal_bstride::Int64
end

abstract type BatchController <: Controller end

"""
    BatchLQR(mechanism, bodyids, eqcids, Q, R, horizon; xd, vd, qd, ωd, Fτd)

Same arguments as `LQR` (reference src/control/lqr.jl:49-66); `xd/qd/...` may be vectors of per-system setpoints.
Gains are synthesised on the GPU (`ccb_lqr_synthesize` = linearsystem + dlqr of lqr.jl:63,141-184).
"""
mutable struct BatchLQR <: BatchController
    mh::MechHandle
    K::Array{Float64,4}        # [12nb][nu][rows][S]  (C order [S][rows][nu][12nb])
    zd::Array{Float64,3}       # [13][nb][S]
    Fτd::Matrix{Float64}       # [nu][S]
    eqcidx::Vector{Int32}
    N::Float64                 # Inf or the horizon in steps (type parameter N of LQR{T,N})
    kexit::Vector{Int32}; iters::Vector{Int32}
end

function BatchLQR(mechanism::Mechanism, bodyids, eqcids, Q, R, horizon; zd::Array{Float64,3}, Fτd = zeros(length(eqcids), size(zd, 3)), device = 0)
    Nb = length(mechanism.bodies)
    @assert length(bodyids) == length(Q) == Nb "Missmatched length for bodies"                  # lqr.jl:59
    @assert length(eqcids) == length(R) == size(Fτd, 1) "Missmatched length for constraints"    # lqr.jl:60
    mh = MechHandle(mechanism)
    Δt = mechanism.Δt
    Qd = Matrix(cat(Q..., dims = (1, 2)) * Δt)                                                   # lqr.jl:18
    Rd = Matrix(cat(R..., dims = (1, 2)) * Δt)                                                   # lqr.jl:19
    N = horizon / Δt                                                                              # lqr.jl:21-27
    Ntemp = N < Inf ? Integer(ceil(horizon / Δt)) : Integer(ceil(10 / Δt))
    N < Inf && (N = Float64(Ntemp))
    eqcidx = Int32[findfirst(e -> getid(e) == id, collect(mechanism.eqconstraints)) - 1 for id in eqcids]
    S = size(zd, 3); nu = length(eqcids); n = 12Nb
    Ku = Array{Float64}(undef, n, nu, Ntemp - 1, S); kexit = zeros(Int32, S); iters = zeros(Int32, S)
    Qrm = permutedims(Qd); Rrm = permutedims(Rd)                                                  # row-major for the ABI
    check(ccall((:ccb_lqr_synthesize, libccb), Cint,
                (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Int32, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Int32,
                 Ptr{Cvoid}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}),
                mh.ptr, device, S, zd, Fτd, eqcidx, nu, Qrm, 0, Rrm, 0, Ntemp, C_NULL, Ku, kexit, iters))
    if N == Inf                                                                                   # lqr.jl:40-43
        any(s -> Ku[:, :, 1, s] != Ku[:, :, 2, s], 1:S) && @info "Riccati recursion did not converge."
        Ku = Ku[:, :, 1:1, :]
    end
    return BatchLQR(mh, Ku, zd, Matrix{Float64}(Fτd), eqcidx, N, kexit, iters)
end

"""
    DeviceLQR(mechanism, bodyids, eqcids, Q, R, horizon; zd, Fτd)

`BatchLQR` whose gain table never leaves the GPU (`ccb_lqr_create`): the right constructor when the gains are only used by
`simulate!`.  `converged[s]` is `Ku[1] == Ku[2]` of lqr.jl:41 (meaningful for `horizon = Inf`).
"""
mutable struct DeviceLQR <: BatchController
    mh::MechHandle
    ptr::Ptr{Cvoid}
    S::Int
    kexit::Vector{Int32}; iters::Vector{Int32}; converged::Vector{Int32}
end

function DeviceLQR(mechanism::Mechanism, bodyids, eqcids, Q, R, horizon; zd::Array{Float64,3}, Fτd = zeros(length(eqcids), size(zd, 3)), device = 0)
    Nb = length(mechanism.bodies)
    @assert length(bodyids) == length(Q) == Nb "Missmatched length for bodies"                  # lqr.jl:59
    @assert length(eqcids) == length(R) == size(Fτd, 1) "Missmatched length for constraints"    # lqr.jl:60
    mh = MechHandle(mechanism)
    Δt = mechanism.Δt
    Qrm = permutedims(Matrix(cat(Q..., dims = (1, 2)) * Δt)); Rrm = permutedims(Matrix(cat(R..., dims = (1, 2)) * Δt))   # lqr.jl:18-19
    inf = horizon / Δt == Inf
    Ntemp = inf ? Integer(ceil(10 / Δt)) : Integer(ceil(horizon / Δt))                           # lqr.jl:21-27
    eqcidx = Int32[findfirst(e -> getid(e) == id, collect(mechanism.eqconstraints)) - 1 for id in eqcids]
    S = size(zd, 3); nu = length(eqcids)
    kexit = zeros(Int32, S); iters = zeros(Int32, S); conv = zeros(Int32, S)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ccb_lqr_create, libccb), Cint,
                (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Int32, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Int32, Int32,
                 Ptr{Cvoid}, Ref{Ptr{Cvoid}}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}),
                mh.ptr, device, S, zd, Fτd, eqcidx, nu, Qrm, 0, Rrm, 0, Ntemp, inf ? 1 : 0, C_NULL, out, kexit, iters, conv))
    inf && any(==(0), conv) && @info "Riccati recursion did not converge."                       # lqr.jl:41
    c = DeviceLQR(mh, out[], S, kexit, iters, conv)
    finalizer(x -> ccall((:ccb_ctrl_destroy, libccb), Cint, (Ptr{Cvoid},), x.ptr), c)
    return c
end
ctrl_handle(lqr::DeviceLQR, B, device) = lqr.ptr
owns_handle(::DeviceLQR) = false
owns_handle(::BatchController) = true

"""
    DeviceTrackingLQR(mechanism, storage::Array{Float64,4}, Fτ::Array{Float64,3}, eqcids, Q, R)

`TrackingLQR(mechanism, storage, Fτ, eqcids, Q, R)` (reference src/control/lqr_tracking.jl:17-43) for S reference trajectories at once:
`storage[:, i, k, s]` = (x, q, v, ω) of body i at step k of trajectory s (the transpose `storage.x[i][k] -> xd[k][i]` of
lqr_tracking.jl:30-37 is the memory layout itself), `Fτ[:, k, s]` the feed-forward inputs.  Gains, desired trajectory and feed-forward
stay on the GPU (`ccb_tracking_create`); `replicate!(ctrl, per)` lets `per` consecutive systems of a batch share one trajectory.
"""
mutable struct DeviceTrackingLQR <: BatchController
    mh::MechHandle
    ptr::Ptr{Cvoid}
    S::Int; N::Int
    kexit::Vector{Int32}; iters::Vector{Int32}
end

function DeviceTrackingLQR(mechanism::Mechanism, storage::Array{Float64,4}, Fτ::Array{Float64,3}, eqcids, Q, R; device = 0)
    Nb = length(mechanism.bodies)
    @assert size(storage, 1) == 13 && size(storage, 2) == Nb
    N, S = size(storage, 3), size(storage, 4)
    mh = MechHandle(mechanism)
    Δt = mechanism.Δt
    Qrm = permutedims(Matrix(cat(Q..., dims = (1, 2)) * Δt)); Rrm = permutedims(Matrix(cat(R..., dims = (1, 2)) * Δt))   # lqr_tracking.jl:22-23
    eqcidx = Int32[findfirst(e -> getid(e) == id, collect(mechanism.eqconstraints)) - 1 for id in eqcids]
    nu = length(eqcids)
    @assert size(Fτ) == (nu, N, S)
    kexit = zeros(Int32, S); iters = zeros(Int32, S)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ccb_tracking_create, libccb), Cint,
                (Ptr{Cvoid}, Int32, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Cvoid}, Ref{Ptr{Cvoid}},
                 Ptr{Int32}, Ptr{Int32}),
                mh.ptr, device, S, N, storage, Fτ, eqcidx, nu, Qrm, Rrm, C_NULL, out, kexit, iters))
    c = DeviceTrackingLQR(mh, out[], S, N, kexit, iters)
    finalizer(x -> ccall((:ccb_ctrl_destroy, libccb), Cint, (Ptr{Cvoid},), x.ptr), c)
    return c
end
replicate!(c::Union{DeviceLQR,DeviceTrackingLQR}, per::Integer) = (check(ccall((:ccb_ctrl_set_replication, libccb), Cint, (Ptr{Cvoid}, Int32), c.ptr, per)); c)
ctrl_handle(c::DeviceTrackingLQR, B, device) = c.ptr
owns_handle(::DeviceTrackingLQR) = false

"""
    BatchPID(mechanism, eqcids, goals::Matrix; P, I, D)

`PID(mechanism, eqcids, goals; P, I, D)` (reference src/control/pid.jl:25-39) with one goal column per system; the controller memory
(`integratederrors`, `lasterrors`, pid.jl:9-10) lives on the device and persists across `simulate!` calls like pid.jl:76.
"""
mutable struct BatchPID <: BatchController
    mh::MechHandle
    ptr::Ptr{Cvoid}
    B::Int
end

function BatchPID(mechanism::Mechanism, eqcids, goals::Matrix{Float64}; P = zeros(length(eqcids)), I = zeros(length(eqcids)), D = zeros(length(eqcids)), device = 0)
    mh = MechHandle(mechanism)
    nu = length(eqcids); B = size(goals, 2)
    @assert size(goals, 1) == nu
    eqcidx = Int32[findfirst(e -> getid(e) == id, collect(mechanism.eqconstraints)) - 1 for id in eqcids]
    Pv, Iv, Dv = Float64.(P), Float64.(I), Float64.(D)
    d = Ref(CtrlDesc(5, nu, 0, B, pointer(eqcidx), C_NULL, 0, C_NULL, 0, C_NULL, 0, C_NULL, 0, pointer(Pv), pointer(Iv), pointer(Dv), pointer(goals), B == 1 ? 0 : nu))
    out = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve eqcidx Pv Iv Dv goals begin
        rc = ccall((:ccb_ctrl_create, libccb), Cint, (Ptr{Cvoid}, Ref{CtrlDesc}, Int32, Ref{Ptr{Cvoid}}), mh.ptr, d, device, out)
    end
    rc == 0 || (rc == -1 && throw(AssertionError("Only 1 DOF joints are supported")))      # pid.jl:20,36
    check(rc)
    c = BatchPID(mh, out[], B)
    finalizer(x -> ccall((:ccb_ctrl_destroy, libccb), Cint, (Ptr{Cvoid},), x.ptr), c)
    return c
end
ctrl_handle(c::BatchPID, B, device) = c.ptr
owns_handle(::BatchPID) = false

function ctrl_handle(lqr::BatchLQR, B, device)
    n, nu, rows, S = size(lqr.K)
    d = Ref(CtrlDesc(lqr.N == Inf ? 2 : 3, nu, lqr.N == Inf ? 0 : Int32(lqr.N), B, pointer(lqr.eqcidx),
                     pointer(lqr.K), S == 1 ? 0 : n * nu * rows, pointer(lqr.zd), size(lqr.zd, 3) == 1 ? 0 : 13 * size(lqr.zd, 2),
                     pointer(lqr.Fτd), size(lqr.Fτd, 2) == 1 ? 0 : nu, C_NULL, 0, C_NULL, C_NULL, C_NULL, C_NULL, 0))
    out = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve lqr check(ccall((:ccb_ctrl_create, libccb), Cint, (Ptr{Cvoid}, Ref{CtrlDesc}, Int32, Ref{Ptr{Cvoid}}), lqr.mh.ptr, d, device, out))
    return out[]
end

"""
    simulate!(batch::BatchStates, tend, controller; record = false)

`simulate!(mechanism, tend, controller; record)` (reference examples/lqr_pendulum.jl:43) for every system of the batch.
Returns the recorded trajectories [13][nb][B][steps] when `record`.
"""
function simulate!(batch::BatchStates, Δt::Real, tend::Real, controller::Union{BatchController,Nothing}; record = false, device = 0)
    steps = Integer(ceil(tend / Δt))
    ch = controller === nothing ? C_NULL : ctrl_handle(controller, batch.B, device)
    storage = record ? Array{Float64}(undef, 13, batch.mech.nb, batch.B, steps) : nothing
    status = zeros(Int32, batch.B)
    rc = ccall((:ccb_simulate, libccb), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int32, Ptr{Cvoid}, Ptr{Float64}, Ptr{Int32}),
               batch.ptr, ch, 1, steps, C_NULL, record ? storage : C_NULL, status)
    (ch == C_NULL || !owns_handle(controller)) || ccall((:ccb_ctrl_destroy, libccb), Cint, (Ptr{Cvoid},), ch)
    check(rc)
    return record ? storage : nothing
end

# ---------------------------------------------------------------------------------------------------------------------------
# Algebraic-Riccati utilities (reference src/util/util.jl) on the device: same names and argument order, matrices may carry a
# trailing batch dimension (n, n, S).  Julia is column-major, the ABI row-major: every matrix goes through `permutedims`.
# ---------------------------------------------------------------------------------------------------------------------------
rowmajor(M::AbstractMatrix) = reshape(permutedims(Matrix{Float64}(M)), size(M, 2), size(M, 1), 1)
rowmajor(M::AbstractArray{<:Real,3}) = permutedims(Array{Float64,3}(M), (2, 1, 3))

function are_solve(kind::Integer, A, B, Q, R; wantX = true, wantK = false, device = 0)
    Ar, Br, Qr, Rr = rowmajor(A), rowmajor(B), rowmajor(Q), rowmajor(R)
    n = size(Ar, 1); m = size(Br, 1); S = maximum(size.((Ar, Br, Qr, Rr), 3))
    rep(M) = size(M, 3) == S ? M : repeat(M, 1, 1, S)
    Ar, Br, Qr, Rr = rep(Ar), rep(Br), rep(Qr), rep(Rr)
    X = wantX ? Array{Float64}(undef, n, n, S) : nothing
    K = wantK ? Array{Float64}(undef, n, m, S) : nothing          # row-major [S][m][n]
    iters = zeros(Int32, S); status = zeros(Int32, S)
    check(ccall((:ccb_are_solve, libccb), Cint,
                (Int32, Int32, Int32, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}),
                device, kind, S, n, m, Ar, Br, Qr, Rr, wantX ? X : C_NULL, wantK ? K : C_NULL, iters, status))
    any(!=(0), status) && @info "Riccati equation: no stabilising solution found for $(count(!=(0), status)) of $S problems."
    single = ndims(A) == 2 && ndims(B) <= 2 && ndims(Q) == 2
    out(M) = single ? permutedims(M[:, :, 1]) : permutedims(M, (2, 1, 3))
    return wantX && wantK ? (out(X), out(K)) : wantX ? out(X) : out(K)
end

dare(A, B, Q, R; kw...) = are_solve(0, A, B, Q, R; kw...)                                   # util.jl:1-20
care(A, B, Q, R; kw...) = are_solve(1, A, B, Q, R; kw...)                                   # util.jl:22-37
lqr(A, B, Q, R; kw...) = are_solve(1, A, B, Q, R; wantX = false, wantK = true, kw...)        # util.jl:39-43
dlqr(A, B, Q, R; kw...) = are_solve(0, A, B, Q, R; wantX = false, wantK = true, kw...)       # util.jl:45-49
dlqr(A, B, Q, R, Δt::Real; kw...) = dlqr(A .* Δt .+ Matrix{Float64}(I, size(A, 1), size(A, 1)), B .* Δt, Q .* Δt, R .* Δt; kw...)   # util.jl:51-59

"give the library's persistent device scratch (kept between synthesis calls) back to the driver"
release_scratch() = check(ccall((:ccb_scratch_release, libccb), Cint, ()))

# ---------------------------------------------------------------------------------------------------------------------------
# Multi-GPU (one Julia process per GPU, e.g. Distributed.jl workers): gain tables that every GPU of the node needs are filled by
# the Riccati sweeps of all ranks together -- each rank's kernel stores its rows into every rank's table over NVLink (CUDA IPC).
#   t, h = gains_alloc(count; device)            my table and its 64-byte IPC handle (send `h` to the other ranks)
#   p = gains_open(h; device)                    map a peer's table
#   synthesize_peers!(mh, zd, Fτd, eqcidx, Qd, Rd, N, tables, s_offset; device)   my shard -> all tables;  then a barrier
# ---------------------------------------------------------------------------------------------------------------------------
function gains_alloc(count::Integer; device = 0)
    p = Ref{Ptr{Float64}}(C_NULL); h = zeros(UInt8, 64)
    check(ccall((:ccb_gains_alloc, libccb), Cint, (Int32, Int64, Ref{Ptr{Float64}}, Ptr{UInt8}), device, count, p, h))
    return p[], h
end
function gains_open(h::Vector{UInt8}; device = 0)
    p = Ref{Ptr{Float64}}(C_NULL)
    check(ccall((:ccb_gains_open, libccb), Cint, (Int32, Ptr{UInt8}, Ref{Ptr{Float64}}), device, h, p))
    return p[]
end
gains_close(p::Ptr{Float64}) = check(ccall((:ccb_gains_close, libccb), Cint, (Ptr{Float64},), p))
gains_free(p::Ptr{Float64}) = check(ccall((:ccb_gains_free, libccb), Cint, (Ptr{Float64},), p))
function synthesize_peers!(mh::MechHandle, zd::Array{Float64,3}, Fτd::Matrix{Float64}, eqcidx::Vector{Int32}, Qd::Matrix{Float64}, Rd::Matrix{Float64},
                           N::Integer, tables::Vector{Ptr{Float64}}, s_offset::Integer; device = 0)
    S = size(zd, 3); nu = length(eqcidx)
    kexit = zeros(Int32, S); iters = zeros(Int32, S)
    check(ccall((:ccb_lqr_synthesize_peers, libccb), Cint,
                (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Int32, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Int32, Ptr{Cvoid},
                 Int32, Ptr{Ptr{Float64}}, Int64, Ptr{Int32}, Ptr{Int32}),
                mh.ptr, device, S, zd, Fτd, eqcidx, nu, permutedims(Qd), 0, permutedims(Rd), 0, N, C_NULL, length(tables), tables, s_offset, kexit, iters))
    return kexit, iters
end
# Optional: keep the rows below the break index (copies, lqr.jl:179-181) off NVLink.  defer_backfill!(true) before synthesize_peers!, then -- after the
# barrier, with the break indices of ALL systems gathered -- gains_backfill!(t, kexit_all, N, nu, n) on every rank's own table, then a second barrier.
defer_backfill!(on::Bool) = check(ccall((:ccb_gains_defer_backfill, libccb), Cint, (Int32,), on ? 1 : 0))
gains_backfill!(t::Ptr{Float64}, kexit::Vector{Int32}, N::Integer, nu::Integer, n::Integer; device = 0) =
    check(ccall((:ccb_gains_backfill, libccb), Cint, (Int32, Ptr{Float64}, Int64, Int32, Int32, Int32, Ptr{Int32}), device, t, length(kexit), N, nu, n, kexit))

end # module
