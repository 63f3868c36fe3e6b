# tools/dump_golden.jl -- pins the oracle (and through it the CUDA path) against the REAL reference.
#
#   julia --project=<env with ConstrainedControl.jl + ConstrainedDynamics ^0.9.1> tools/dump_golden.jl [outdir]
#
# Runs the unmodified reference on deterministic (literal, rand-free) versions of its own examples and dumps, per case,
#   A, Bu, Bλ, G            linearsystem(mechanism, xd, vd, qd, ωd, Fτd, bodyids, eqcids)      (call site src/control/lqr.jl:63)
#   K                       lqr.K  ->  [rows][nu][12 Nb]                                        (src/control/lqr.jl:39-46, lqr_tracking.jl:40)
#   storage                 simulate!(mech, tend, controller, record = true)  -> [steps][Nb][13] = (x, q, v, ω)
#   dfs_list, ids           mechanism.system.dfs_list, getid.(bodies), getid.(eqconstraints)    (bit-exact contract, SURVEY.md 8(c))
# as raw little-endian binaries + a text manifest (no package beyond the reference's own dependencies is needed).
# tests/test_reference_golden.py loads <repo>/tests/golden/julia/manifest.txt when it exists and compares oracle and GPU with it;
# until somebody with a Julia toolchain runs this file, parity with the upstream implementation stays UNPINNED (DESIGN.md section 2).
#
# NOT RUN IN THE BUILD CONTAINER (no julia, no network).  The cases mirror examples/*.jl line by line; initial conditions that the
# examples draw with rand() are replaced by the literals below, which tests/test_reference_golden.py repeats.
using ConstrainedDynamics
using ConstrainedControl
using LinearAlgebra
using StaticArrays

const OUT = length(ARGS) >= 1 ? ARGS[1] : joinpath(@__DIR__, "..", "tests", "golden", "julia")
mkpath(OUT)
const MANIFEST = IOBuffer()

# column-major Julia arrays are written as they lie; the manifest records the Julia shape and the loader reshapes with order='F'
function put(case::String, name::String, a::AbstractArray{Float64})
    open(joinpath(OUT, "$(case)__$(name).f64"), "w") do io; write(io, Array(a)); end
    println(MANIFEST, "$case $name f64 ", join(size(a), "x"))
end
function put(case::String, name::String, a::AbstractArray{<:Integer})
    open(joinpath(OUT, "$(case)__$(name).i64"), "w") do io; write(io, Array{Int64}(a)); end
    println(MANIFEST, "$case $name i64 ", join(size(a), "x"))
end

# scalar-first quaternion; `Quaternion` is the type the reference's examples construct (`Quaternion(RotX(ϕ))`, examples/lqr_pendulum.jl:29) and whose
# vector part the control laws read with `imag` (src/control/lqr.jl:102); its scalar field is `s` or `w` depending on the ConstrainedDynamics release
quatvec(q) = [hasproperty(q, :s) ? q.s : q.w; imag(q)...]

function storage_array(storage, Nb)
    N = length(storage.x[1])
    S = zeros(13, Nb, N)
    for k = 1:N, i = 1:Nb
        S[1:3, i, k] = storage.x[i][k]
        S[4:7, i, k] = quatvec(storage.q[i][k])
        S[8:10, i, k] = storage.v[i][k]
        S[11:13, i, k] = storage.ω[i][k]
    end
    return S                              # Julia (13, Nb, N)  ==  numpy [N][Nb][13]
end

function gains_array(K)                  # Vector{Vector{SMatrix{1,NK}}} -> (NK, nu, rows) == numpy [rows][nu][NK]
    rows, nu, NK = length(K), length(K[1]), length(K[1][1])
    A = zeros(NK, nu, rows)
    for k = 1:rows, i = 1:nu
        A[:, i, k] = vec(K[k][i])
    end
    return A
end

function dump_graph(case, mech)
    put(case, "dfs_list", collect(Int64, mech.system.dfs_list))
    put(case, "body_ids", collect(Int64, getid.(mech.bodies)))
    put(case, "eqc_ids", collect(Int64, getid.(mech.eqconstraints)))
    put(case, "parents", collect(Int64, [isempty(p) ? 0 : p[1] for p in mech.system.parents]))
end

function dump_lqr_case(case, mech, bodies, eqcs_ctrl, Q, R, horizon, tend; xd, qd = [one(Quaternion{Float64}) for _ in bodies],
                       Fτd = [SA[0.0] for _ in eqcs_ctrl])
    Nb = length(bodies)
    vd = [SA[0.0; 0; 0] for _ = 1:Nb]; ωd = [SA[0.0; 0; 0] for _ = 1:Nb]
    dump_graph(case, mech)
    x0 = zeros(13, Nb)
    for (i, b) in enumerate(bodies)
        x0[:, i] = [b.state.xc; quatvec(b.state.qc); b.state.vc; b.state.ωc]
    end
    put(case, "state0", x0)
    A, Bu, Bλ, G = ConstrainedDynamics.linearsystem(mech, xd, vd, qd, ωd, Fτd, getid.(bodies), getid.(eqcs_ctrl))     # lqr.jl:63
    put(case, "A", Matrix(A)); put(case, "Bu", Matrix(Bu)); put(case, "Bl", Matrix(Bλ)); put(case, "G", Matrix(G))
    lqr = LQR(mech, getid.(bodies), getid.(eqcs_ctrl), Q, R, horizon, xd = xd, qd = qd, Fτd = Fτd)                    # lqr.jl:49-66
    put(case, "K", gains_array(lqr.K))
    storage = simulate!(mech, tend, lqr, record = true)                                                                # examples/lqr_pendulum.jl:43
    put(case, "storage", storage_array(storage, Nb))
end

ex = [1.0; 0.0; 0.0]; ey = [0.0; 1.0; 0.0]
p2 = [0.0; 0.0; 0.5]

# ---- C1 examples/lqr_pendulum.jl (already deterministic)
let
    origin = Origin{Float64}()
    link1 = Box(0.1, 0.1, 1.0, 1.0)
    joint = EqualityConstraint(Revolute(origin, link1, ex; p2 = p2))
    mech = Mechanism(origin, [link1], [joint])
    setPosition!(origin, link1, p2 = p2, Δq = Quaternion(RotX(pi - 0.4)))
    Q = [diagm(ones(12)) * 0.0]; Q[1][7, 7] = 1000.0; Q[1][10, 10] = 100.0
    dump_lqr_case("c1_pendulum", mech, [link1], [joint], Q, [ones(1, 1)], Inf, 10.0; xd = [[0; 0.0; 0.5]], qd = [Quaternion(RotX(pi))])
end

# ---- C2 examples/lqr_cartpole.jl:8-41 (its own initial conditions y0 = 0.5, θ0 = 0.2), horizons 10 (the example) and 20 (BASELINE)
for (tag, horizon) in (("c2_cartpole_h10", 10.0), ("c2_cartpole_h20", 20.0))
    origin = Origin{Float64}()
    cart = Box(0.1, 0.5, 0.1, 0.5)
    pole = Box(0.1, 0.1, 1.0, 1.0)
    joint1 = EqualityConstraint(Prismatic(origin, cart, ey))
    joint2 = EqualityConstraint(Revolute(cart, pole, ex; p2 = -p2))
    mech = Mechanism(origin, [cart; pole], [joint1; joint2], g = -9.81)
    setPosition!(origin, cart, Δx = [0; 0.5; 0])
    setPosition!(cart, pole, p2 = -p2, Δq = Quaternion(RotX(0.2)))
    Q = [diagm(ones(12)) * 1.0 for i = 1:2]            # examples/lqr_cartpole.jl:38-39
    dump_lqr_case(tag, mech, [cart; pole], [joint1], Q, [ones(1, 1)], horizon, horizon; xd = [[0; 0; 0.0], [0; 0; 0.5]])
end

# ---- C3 examples/lqr_cartpole_n_pendulum.jl with N = 3 links, ϕinit = [0.02, 0.01, 0.005], yinit = 0.2
let
    N = 3; ϕinit = [0.02, 0.01, 0.005]; yinit = 0.2
    origin = Origin{Float64}()
    cart = Box(0.1, 0.5, 0.1, 0.5)
    bodies = [cart; [Box(0.1, 0.1, 1.0, 1.0) for i = 1:N]]
    constraints = [EqualityConstraint(Prismatic(origin, cart, ey)); EqualityConstraint(Revolute(cart, bodies[2], ex; p2 = -p2))]
    constraints = [constraints; [EqualityConstraint(Revolute(bodies[i], bodies[i + 1], ex; p1 = p2, p2 = -p2)) for i = 2:N]]
    mech = Mechanism(origin, bodies, constraints, g = -9.81)
    setPosition!(origin, cart, Δx = [0; yinit; 0])
    setPosition!(cart, bodies[2], p2 = -p2, Δq = Quaternion(RotX(ϕinit[1])))
    for i = 2:N
        setPosition!(bodies[i], bodies[i + 1], p1 = p2, p2 = -p2, Δq = Quaternion(RotX(ϕinit[i])))
    end
    xd = [[[0; 0; 0.0]]; [[0; 0; i - 1 + 0.5] for i = 1:N]]
    dump_lqr_case("c3_npendulum3", mech, bodies, [constraints[1]], [diagm(ones(12)) for i = 1:N + 1], [ones(1, 1)], 10.0, 10.0; xd = xd)
end

# ---- C4 examples/lqr_sawyer.jl (zero configuration set point), started 0.01 rad off on every joint
let
    path = joinpath(@__DIR__, "..", "..", "reference", "examples", "examples_files", "sawyer_arm.urdf")
    isfile(path) || (path = "examples/examples_files/sawyer_arm.urdf")
    mech = Mechanism(path, floating = false, g = 0.0)
    for j = 0:6
        setPosition!(mech, mech.eqconstraints["right_j$j"], [0.0])
    end
    bodies = collect(mech.bodies); eqcs = collect(mech.eqconstraints)
    xd = [body.state.xc for body in bodies]; qd = [body.state.qc for body in bodies]
    for j = 0:6
        setPosition!(mech, mech.eqconstraints["right_j$j"], [0.01])
    end
    put("c4_sawyer", "body_names_order", collect(Int64, getid.(bodies)))
    dump_lqr_case("c4_sawyer", mech, bodies, eqcs, [diagm(ones(12)) * 1000.0 for i = 1:7], [ones(1, 1) for i = 1:7], 20.0, 2.0; xd = xd, qd = qd)
end

# ---- C5 examples/pid_doublependulum.jl (deterministic)
let
    origin = Origin{Float64}()
    link1 = Box(0.1, 0.1, 1.0, 1.0); link2 = deepcopy(link1)
    joint1 = EqualityConstraint(Revolute(origin, link1, ex; p2 = p2))
    joint2 = EqualityConstraint(Revolute(link1, link2, ex; p1 = -p2, p2 = p2))
    mech = Mechanism(origin, [link1; link2], [joint1; joint2])
    setPosition!(origin, link1, p2 = p2, Δq = Quaternion(RotX(0.0)))
    setPosition!(link1, link2, p1 = -p2, p2 = p2, Δq = Quaternion(RotX(0.0)))
    dump_graph("c5_pid_doublependulum", mech)
    pid = PID(mech, getfield.([joint1; joint2], :id), [pi / 2; -pi / 4], P = [10.0; 10.0], I = [10.0; 10], D = [5.0; 5.0])
    storage = simulate!(mech, 10.0, pid, record = true)
    put("c5_pid_doublependulum", "storage", storage_array(storage, 2))
end

# ---- TrackingLQR: examples/trackingLQR_triple_cartpole.jl:1-117 without the noise terms (stock control_trackinglqr!, lqr_tracking.jl:46-71)
let
    Ufile = joinpath(@__DIR__, "..", "tests", "golden", "triple_cartpole_U.txt")      # one value per line; written by tests/golden/make_golden.py
    U = parse.(Float64, readlines(Ufile))
    origin = Origin{Float64}()
    cart = Box(0.1, 0.5, 0.1, 0.5)
    pole1 = Box(0.1, 0.1, 1.0, 1.0); pole2 = deepcopy(pole1); pole3 = deepcopy(pole1)
    joint1 = EqualityConstraint(Prismatic(origin, cart, ey))
    joint2 = EqualityConstraint(Revolute(cart, pole1, ex; p2 = p2))
    joint3 = EqualityConstraint(Revolute(pole1, pole2, ex; p1 = -p2, p2 = p2))
    joint4 = EqualityConstraint(Revolute(pole2, pole3, ex; p1 = -p2, p2 = p2))
    links = [cart; pole1; pole2; pole3]
    mech = Mechanism(origin, links, [joint1; joint2; joint3; joint4], g = -9.81, Δt = 0.01)
    function reset!()
        setPosition!(origin, cart, Δx = [0; 0.0; 0])
        setPosition!(cart, pole1, p2 = p2, Δq = Quaternion(RotX(0.0)))
        setPosition!(pole1, pole2, p1 = -p2, p2 = p2, Δq = Quaternion(RotX(0.0)))
        setPosition!(pole2, pole3, p1 = -p2, p2 = p2, Δq = Quaternion(RotX(0.0)))
        for b in links; setVelocity!(b); end
    end
    reset!()
    dump_graph("tracking_triple_cartpole", mech)
    control!(mechanism, k) = setForce!(mechanism, joint1, [U[k]])
    storage0 = Storage{Float64}(Base.OneTo(1000), 4)
    simulate!(mech, storage0, control!, record = true)
    put("tracking_triple_cartpole", "storage0", storage_array(storage0, 4))
    Q = [diagm(ones(12)) * 0.0 for i = 1:4]
    Q[1][2, 2] = 10; Q[1][5, 5] = 1
    for i = 2:4; Q[i][7, 7] = 40; Q[i][10, 10] = 1; end
    lqr = TrackingLQR(mech, storage0, [[[U[k]]] for k = 1:1000], [joint1.id], Q, [ones(1, 1) * 0.1])
    put("tracking_triple_cartpole", "K", gains_array(lqr.K))
    reset!()
    setPosition!(cart, pole1, p2 = p2, Δq = Quaternion(RotX(0.01)))                # a perturbed start that the tracking controller pulls back
    put("tracking_triple_cartpole", "state0_tracked", hcat([[b.state.xc; quatvec(b.state.qc); b.state.vc; b.state.ωc] for b in links]...))
    storage = Storage{Float64}(Base.OneTo(1000), 4)
    simulate!(mech, storage, lqr, record = true)
    put("tracking_triple_cartpole", "storage_tracked", storage_array(storage, 4))
end

open(joinpath(OUT, "manifest.txt"), "w") do io; write(io, String(take!(MANIFEST))); end
println("wrote ", OUT)
