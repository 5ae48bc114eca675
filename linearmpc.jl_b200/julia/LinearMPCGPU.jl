# LinearMPCGPU.jl -- the thin glue a LinearMPC.jl maintainer would add as src/gpu/.
# It binds liblmpcgpu.so (include/lmpc_gpu.h) with ccall and keeps the reference's API shape:
#   setup_gpu!(mpc)                          after setup!(mpc)          (src/setup.jl:7-29)
#   compute_control(gmpc, X; r=R, d=D, uprev=U)  X::AbstractMatrix => batch (src/utils.jl:43-51)
#   solve(gmpc, Θ)                                                        (src/utils.jl:268-283)
#   Simulation-style closed loops: simulate(gmpc, X0; r, d, N, warm_start) (src/simulation.jl:93-113)
#   compute_control(gempc, X; ...)            explicit controllers          (src/utils.jl:53-60)
# NOT executed in the build container (no Julia there); every call below is a 1:1 ccall of a
# function that the Python mirror (linearmpc.jl_b200/host.py) exercises in the tests.
module LinearMPCGPU

using LinearMPC, LinearAlgebra
import ParametricDAQP

const liblmpcgpu = get(ENV, "LMPC_GPU_LIB", "liblmpcgpu.so")

struct LmpcProblem            # mirrors lmpc_problem in include/lmpc_gpu.h
    n::Int32; m::Int32; ms::Int32; nth::Int32; nu::Int32; nx::Int32
    H::Ptr{Float64}; f::Ptr{Float64}; f_theta::Ptr{Float64}; A::Ptr{Float64}
    bu::Ptr{Float64}; bl::Ptr{Float64}; W::Ptr{Float64}; senses::Ptr{Cint}
    K::Ptr{Float64}; soft_weight::Float64
    has_binaries::Int32; is_symmetric::Int32; n_break_points::Int32
end

mutable struct LmpcSettings    # mirrors lmpc_settings in include/lmpc_gpu.h (DAQP settings, docs/src/manual/solver.md:49-56)
    primal_tol::Float64; dual_tol::Float64; zero_tol::Float64; progress_tol::Float64; fval_bound::Float64
    cycle_tol::Int32; iter_limit::Int32; kernel::Int32; certified_iterations::Int32
    LmpcSettings() = (s = new(); ccall((:lmpc_gpu_default_settings, liblmpcgpu), Cvoid, (Ref{LmpcSettings},), s); s)
end

mutable struct GpuMPC
    handle::Ptr{Cvoid}
    mpc::LinearMPC.MPC
    nth::Int; nu::Int; m::Int
    multi::Bool            # handle is an lmpc_multi (one replica per device, host batches only)
end

"""
    pinned(T, dims...) -> Array{T}

Array in page-locked host memory (`lmpc_gpu_host_alloc`): batches allocated this way are DMA'd in
place (9.5e8 invpend solves/s end to end per GPU); ordinary Julia arrays are staged through the
library's pinned ring (about 7e8 when the arrays are reused between calls).  Freed by a finalizer.
"""
function pinned(::Type{T}, dims::Integer...) where {T}
    p = ccall((:lmpc_gpu_host_alloc, liblmpcgpu), Ptr{Cvoid}, (Csize_t,), max(prod(dims) * sizeof(T), 8))
    p == C_NULL && error("lmpc_gpu_host_alloc failed: $(lasterror())")
    A = unsafe_wrap(Array, Ptr{T}(p), dims; own=false)
    finalizer(_ -> ccall((:lmpc_gpu_host_free, liblmpcgpu), Cvoid, (Ptr{Cvoid},), p), A)
    return A
end

lasterror() = unsafe_string(ccall((:lmpc_gpu_last_error, liblmpcgpu), Cstring, ()))
check(rc) = rc == 0 || error("liblmpcgpu error $rc: $(lasterror())")

"""
    setup_gpu!(mpc; device=0) -> GpuMPC

Uploads the condensed problem once (what `DAQP.setup` + `qp2ldp` do on the host,
src/setup.jl:11-13, src/codegen.jl:239-280).  Refuses hybrid / AVI / hierarchical problems.
"""
function setup_gpu!(mpc::LinearMPC.MPC; device::Integer=0, devices=nothing, certificate=nothing)
    mpc.mpqp_issetup || LinearMPC.setup!(mpc)
    # certificate: a CertificationResult of certify(mpc; range) (src/certify.jl:18-29).  Its proven iteration
    # bound becomes the hard iteration cap and sizes the on-chip working-set capacity (lmpc_settings).
    st = LmpcSettings()
    isnothing(certificate) || (st.certified_iterations = Int32(certificate.max_iterations))
    q = mpc.mpQP
    n, m = size(q.H, 1), length(q.bu)
    nth = size(q.f_theta, 2)
    K = iszero(mpc.K) ? C_NULL : pointer(mpc.K)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve q mpc begin
        p = LmpcProblem(n, m, m - size(q.A, 1), nth, mpc.model.nu, mpc.model.nx,
                        pointer(q.H), pointer(q.f), pointer(q.f_theta),
                        isempty(q.A) ? C_NULL : pointer(q.A), pointer(q.bu), pointer(q.bl),
                        pointer(q.W), pointer(q.senses), K, mpc.settings.soft_weight,
                        q.has_binaries, q.is_symmetric, length(q.break_points))
        if isnothing(devices)
            check(ccall((:lmpc_gpu_create, liblmpcgpu), Cint,
                        (Ref{LmpcProblem}, Ref{LmpcSettings}, Cint, Ref{Ptr{Cvoid}}), p, st, device, h))
        else    # one process, several GPUs: contiguous theta ranges, no inter-GPU traffic
            dv = Int32.(collect(devices))
            check(ccall((:lmpc_gpu_create_multi, liblmpcgpu), Cint,
                        (Ref{LmpcProblem}, Ref{LmpcSettings}, Ptr{Int32}, Int32, Ref{Ptr{Cvoid}}),
                        p, st, dv, length(dv), h))
        end
    end
    g = GpuMPC(h[], mpc, nth, mpc.model.nu, m, !isnothing(devices))
    finalizer(g) do x
        x.multi ? ccall((:lmpc_gpu_destroy_multi, liblmpcgpu), Cint, (Ptr{Cvoid},), x.handle) :
                  ccall((:lmpc_gpu_destroy, liblmpcgpu), Cint, (Ptr{Cvoid},), x.handle)
    end
    return g
end

"""
    solve(g, Θ) -> (U, exitflag, iters)     Θ is nth × N (one parameter per column)
"""
function solve(g::GpuMPC, Θ::Matrix{Float64})
    N = size(Θ, 2)
    U = Matrix{Float64}(undef, g.nu, N); flag = Vector{Int32}(undef, N); it = Vector{Int32}(undef, N)
    if g.multi
        check(ccall((:lmpc_gpu_solve_batch_multi, liblmpcgpu), Cint,
                    (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{UInt64}, Ptr{UInt64}),
                    g.handle, Θ, N, U, flag, it, C_NULL, C_NULL))
    else
        check(ccall((:lmpc_gpu_solve_batch, liblmpcgpu), Cint,
                    (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{UInt64}, Ptr{UInt64}),
                    g.handle, Θ, N, U, flag, it, C_NULL, C_NULL))
    end
    return U, flag, it
end

"""
    solve_warm!(g, Θ, active_upper, active_lower) -> (U, exitflag, iters)

Warm-started batch (DAQP keeps its workspace between `solve` calls, src/utils.jl:278-282): every
column starts from the working set in `active_upper` / `active_lower` (`cld(m, 64)` words per
column, e.g. what the previous call left there) and the two arrays are overwritten with the final sets.
"""
function solve_warm!(g::GpuMPC, Θ::Matrix{Float64}, active_upper::Matrix{UInt64}, active_lower::Matrix{UInt64})
    g.multi && error("solve_warm!: single-device backends only")
    N = size(Θ, 2)
    U = Matrix{Float64}(undef, g.nu, N); flag = Vector{Int32}(undef, N); it = Vector{Int32}(undef, N)
    check(ccall((:lmpc_gpu_solve_batch_warm, liblmpcgpu), Cint,
                (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{UInt64}, Ptr{UInt64}),
                g.handle, Θ, N, U, flag, it, active_upper, active_lower))
    return U, flag, it
end

"""
    compute_control(g, X; r, d, uprev, p, check=true) -> nu × N

Batched `compute_control` (src/utils.jl:43-51): X is nx × N; r, d, uprev, p are vectors (broadcast)
or matrices with N columns.  θ = [x; r; d; uprev; p] exactly as `form_parameter` (src/explicit.jl:54-63);
p holds the generalized parameters (np rows, constant over the horizon: `parameter_preview` off).
"""
function LinearMPC.compute_control(g::GpuMPC, X::AbstractMatrix; r=nothing, d=nothing, uprev=nothing, p=nothing, check=true)
    nx, nr, nd, nup, np = LinearMPC.get_parameter_dims(g.mpc)
    (np == 0 || !g.mpc.settings.parameter_preview) || error("parameter_preview is not supported by the GPU backend")
    N = size(X, 2)
    col(v, k) = v isa AbstractMatrix ? v[1:k, :] : repeat(isnothing(v) ? zeros(k) : v[1:k], 1, N)
    Θ = Matrix{Float64}([X; col(r, nr); col(d, nd); col(isnothing(uprev) ? g.mpc.uprev : uprev, nup); col(p, np)])
    U, flag, _ = solve(g, Θ)
    check && @assert all(flag .>= 1)
    g.mpc.uprev .= U[:, end]
    return U
end

"""
    compute_control_shared(g, X; r, d, uprev) -> (U, n_failed)

Monte-Carlo form of `compute_control` (lmpc_gpu_compute_control_batch_bcast): X is nx × N, `r`, `d`, `uprev`
are single vectors shared by all columns and travel to the GPU once; the exit flags stay on the device and
only the number of columns with exitflag < 1 comes back (what `@assert exitflag ≥ 1`, src/utils.jl:46, checks).
"""
function compute_control_shared(g::GpuMPC, X::Matrix{Float64}; r=nothing, d=nothing, uprev=nothing)
    g.multi && error("compute_control_shared: single-device backends only")
    nx, nr, nd, nup, _ = LinearMPC.get_parameter_dims(g.mpc)
    N = size(X, 2)
    vec_or(v, k) = k == 0 ? Float64[] : Vector{Float64}(isnothing(v) ? zeros(k) : v[1:k])
    rv, dv, uv = vec_or(r, nr), vec_or(d, nd), vec_or(isnothing(uprev) ? g.mpc.uprev : uprev, nup)
    U = Matrix{Float64}(undef, g.nu, N)
    nfail = Ref{Int64}(0)
    check(ccall((:lmpc_gpu_compute_control_batch_bcast, liblmpcgpu), Cint,
                (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32, Int32, Int32,
                 Int64, Int32, Ptr{Int32}, Ref{Int64}),
                g.handle, U, X, nr == 0 ? C_NULL : pointer(rv), nd == 0 ? C_NULL : pointer(dv),
                nup == 0 ? C_NULL : pointer(uv), nr, nd, nup, N, 7, C_NULL, nfail))
    return U, nfail[]
end

"""
    simulate(g, X0; r, d, N, warm_start=true) -> (Xend, Ulast, exitflag)

N_traj independent copies of the `Simulation` loop (src/simulation.jl:93-113) with the model's
linear dynamics (src/model.jl:70), N steps each, entirely on the GPU.  `r` / `d` follow the
Scenario convention (src/simulation.jl:70-80): a vector is held constant, a matrix is a trajectory
with one column per step, its last column held when it is shorter than the horizon; the same
scenario drives every closed loop.
"""
function simulate(g::GpuMPC, X0::Matrix{Float64}; r=nothing, d=nothing, N::Integer=100, warm_start::Bool=true)
    g.multi && return simulate_multi(g, X0; r, d, N, warm_start)
    mdl = g.mpc.model
    nx, nr, nd, nup, _ = LinearMPC.get_parameter_dims(g.mpc)
    M = size(X0, 2)
    function traj(v, k)              # k x N, last column held
        (k == 0 || isnothing(v)) && return C_NULL
        V = v isa AbstractMatrix ? Matrix{Float64}(v) : reshape(Vector{Float64}(v), k, 1)
        return hcat(V[:, 1:min(N, size(V, 2))], repeat(V[:, end], 1, max(0, N - size(V, 2))))
    end
    Rt, Dt = traj(r, nr), traj(d, nd)
    Xend = similar(X0); Ulast = Matrix{Float64}(undef, g.nu, M); flag = Vector{Int32}(undef, M)
    check(ccall((:lmpc_gpu_simulate_batch_traj, liblmpcgpu), Cint,
                (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32, Int32, Int32, Int32,
                 Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                 Int64, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}),
                g.handle, mdl.F, mdl.G, nd == 0 ? C_NULL : mdl.Gd, nx, nr, nd, nup,
                X0, C_NULL, C_NULL, C_NULL, Rt, Dt, M, N, warm_start, Xend, Ulast, C_NULL, flag, C_NULL))
    return Xend, Ulast, flag
end

# every GPU of the box from one process (lmpc_gpu_simulate_batch_multi): contiguous ranges of X0's columns
function simulate_multi(g::GpuMPC, X0::Matrix{Float64}; r=nothing, d=nothing, N::Integer=100, warm_start::Bool=true)
    mdl = g.mpc.model
    nx, nr, nd, nup, _ = LinearMPC.get_parameter_dims(g.mpc)
    M = size(X0, 2)
    hold(v, k) = (k == 0 || isnothing(v)) ? nothing :
        (V = v isa AbstractMatrix ? Matrix{Float64}(v) : reshape(Vector{Float64}(v), k, 1);
         hcat(V[:, 1:min(N, size(V, 2))], repeat(V[:, end], 1, max(0, N - size(V, 2)))))
    Rt, Dt = hold(r, nr), hold(d, nd)
    p(x) = isnothing(x) ? Ptr{Float64}(C_NULL) : pointer(x)
    Xend = similar(X0); Ulast = Matrix{Float64}(undef, g.nu, M); flag = Vector{Int32}(undef, M)
    GC.@preserve Rt Dt begin
        ext = Ref(LmpcSimExt(p(Rt), p(Dt), C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, 0, C_NULL, 0, 0))
        check(ccall((:lmpc_gpu_simulate_batch_multi, liblmpcgpu), Cint,
                    (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32, Int32, Int32, Int32,
                     Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{LmpcSimExt},
                     Int64, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}),
                    g.handle, mdl.F, mdl.G, nd == 0 ? C_NULL : mdl.Gd, nx, nr, nd, nup,
                    X0, C_NULL, C_NULL, C_NULL, ext, M, N, warm_start, Xend, Ulast, C_NULL, flag, C_NULL))
    end
    return Xend, Ulast, flag
end

# mirrors lmpc_sim_ext (include/lmpc_gpu.h)
struct LmpcSimExt
    Rtraj::Ptr{Float64}; Dtraj::Ptr{Float64}
    Ft::Ptr{Float64}; Gt::Ptr{Float64}; Gdt::Ptr{Float64}
    C::Ptr{Float64}; Dd::Ptr{Float64}; K::Ptr{Float64}
    Vtraj::Ptr{Float64}
    ny::Int32
    Xhat_end::Ptr{Float64}
    rstride::Int32; dstride::Int32     # preview windows: step t reads its nr / nd entries at t * stride (0 = nr / nd)
end

# mirrors lmpc_sim_stepped (include/lmpc_gpu.h)
struct LmpcSimStepped
    dynamics::Ptr{Cvoid}; user::Ptr{Cvoid}
    Ft::Ptr{Float64}; Gt::Ptr{Float64}; Gdt::Ptr{Float64}; ft::Ptr{Float64}
    ny::Int32
    Cm::Ptr{Float64}; Ddm::Ptr{Float64}; hm::Ptr{Float64}; Vtraj::Ptr{Float64}
    nxo::Int32; ndm::Int32
    Fo::Ptr{Float64}; Go::Ptr{Float64}; Gdo::Ptr{Float64}; fo::Ptr{Float64}
    Co::Ptr{Float64}; Ddo::Ptr{Float64}; ho::Ptr{Float64}; Ko::Ptr{Float64}
    Rtraj::Ptr{Float64}; Dtraj::Ptr{Float64}
    Xhat_end::Ptr{Float64}
end

# Layout self-check (lmpc_gpu_abi_layout, include/lmpc_gpu.h): the two structs above are hand-written
# mirrors, so compare sizeof / fieldoffset with what the loaded library was compiled with and refuse
# to run on a mismatch instead of letting C read past a Julia object.
function check_abi()
    n = ccall((:lmpc_gpu_abi_layout, liblmpcgpu), Cint, (Ptr{Int32}, Int32), C_NULL, 0)
    v = Vector{Int32}(undef, n)
    ccall((:lmpc_gpu_abi_layout, liblmpcgpu), Cint, (Ptr{Int32}, Int32), v, n)
    mirrors = Dict(1 => LmpcProblem, 2 => LmpcSettings, 3 => LmpcSimExt, 6 => LmpcSimStepped)
    i = 1
    while i <= n
        id, sz, nf = v[i], v[i+1], v[i+2]
        offs = v[i+3:i+2+nf]
        if haskey(mirrors, id)
            T = mirrors[id]
            (sizeof(T) == sz && fieldcount(T) == nf && all(fieldoffset(T, k) == offs[k] for k in 1:nf)) ||
                error("LinearMPCGPU: struct $(T) does not match the loaded liblmpcgpu (size $(sizeof(T)) vs $sz)")
        end
        i += 3 + nf
    end
    return true
end
__init__() = check_abi()

"""
    simulate_observed(g, X0; N, r, d, plant=(F,G,Gd), noise) -> (Xend, Xhat_end, Ulast, exitflag)

The closed loop of src/simulation.jl:92-110 with `mpc.state_observer` (a `KalmanFilter`,
src/observer.jl:1-11) in the loop: estimate initialised at X0, `correct!` / `compute_control` on the
estimate / `predict!` every step, true plant `plant` (default: the model), measurement noise `noise`
(ny × T, shared by all closed loops).
"""
function simulate_observed(g::GpuMPC, X0::Matrix{Float64}; N::Integer=100, r=nothing, d=nothing,
                           plant=nothing, noise=nothing, warm_start::Bool=true)
    mdl = g.mpc.model; kf = g.mpc.state_observer
    nx, nr, nd, nup, _ = LinearMPC.get_parameter_dims(g.mpc)
    M = size(X0, 2)
    hold(v, k) = (k == 0 || isnothing(v)) ? nothing :
        (V = v isa AbstractMatrix ? Matrix{Float64}(v) : reshape(Vector{Float64}(v), k, 1);
         hcat(V[:, 1:min(N, size(V, 2))], repeat(V[:, end], 1, max(0, N - size(V, 2)))))
    # Controllers condensed with reference_preview / disturbance_preview read a WINDOW of Np columns per
    # step (get_preview, src/simulation.jl:102-104,128-134): flat trajectory of N + Np - 1 columns with the
    # column N held (rs[:, min(k+i, N)]), window start advancing by one column (= stride) per step.
    Np = g.mpc.Np
    rprev = g.mpc.settings.reference_preview && nr > 0
    dprev = g.mpc.settings.disturbance_preview && nd > 0
    nyr, ndb = rprev ? nr ÷ Np : nr, dprev ? nd ÷ Np : nd
    # r: get_preview(rs, k, Np) reads columns k+1 .. k+Np; d: get_preview(ds, k-1, Np) reads k .. k+Np-1
    window(V, first) = isnothing(V) ? nothing : hcat(V[:, 1+first:end], repeat(V[:, end], 1, Np - 1 + first))
    Rt, Dt = hold(r, nyr), hold(d, ndb)
    rprev && (Rt = window(Rt, 1)); dprev && (Dt = window(Dt, 0))
    rstride, dstride = rprev ? nyr : 0, dprev ? ndb : 0
    ny = size(kf.C, 1)
    Vt = hold(noise, ny)
    Ft, Gt, Gdt = isnothing(plant) ? (nothing, nothing, nothing) : Matrix{Float64}.(plant)
    p(x) = isnothing(x) ? Ptr{Float64}(C_NULL) : pointer(x)
    Xend = similar(X0); Xhat = similar(X0); Ulast = Matrix{Float64}(undef, g.nu, M); flag = Vector{Int32}(undef, M)
    Cm, Ddm, Km = Matrix{Float64}(kf.C), Matrix{Float64}(kf.Dd), Matrix{Float64}(kf.K)
    GC.@preserve Rt Dt Vt Ft Gt Gdt Cm Ddm Km Xhat begin
        ext = Ref(LmpcSimExt(p(Rt), p(Dt), p(Ft), p(Gt), p(Gdt), pointer(Cm), nd == 0 ? C_NULL : pointer(Ddm),
                             pointer(Km), p(Vt), ny, pointer(Xhat), rstride, dstride))
        check(ccall((:lmpc_gpu_simulate_batch_ex, liblmpcgpu), Cint,
                    (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32, Int32, Int32, Int32,
                     Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{LmpcSimExt},
                     Int64, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}),
                    g.handle, mdl.F, mdl.G, nd == 0 ? C_NULL : mdl.Gd, nx, nr, nd, nup,
                    X0, C_NULL, C_NULL, C_NULL, ext, M, N, warm_start, Xend, Ulast, C_NULL, flag, C_NULL))
    end
    return Xend, Xhat, Ulast, flag
end

"""
    simulate_stepped(g, X0; N, r, d, dynamics=nothing, plant=nothing, noise=nothing, warm_start=true)
        -> (Xend, Xhat, Ulast, exitflag)

The Simulation loop (src/simulation.jl:93-113) stepped from the host for what the fused loops cannot express:
a user-defined true plant and the controller's own `state_observer` with its full state, i.e. a `KalmanFilter`
or the augmented estimator of an `OffsetFreeObserver` (src/observer.jl:12-47, 203-222; the estimated disturbances
are appended to the measured ones in θ).  `dynamics` is a C-callable
`(user, stream, N, k, X, U, D, Xnext) -> Cint` (e.g. `@cfunction` around a CUDA.jl kernel launch on `stream`)
that steps all N plants on the device; without it the affine plant `plant = (Ft, Gt, Gdt, ft)` or the model runs.
"""
function simulate_stepped(g::GpuMPC, X0::Matrix{Float64}; N::Integer=100, r=nothing, d=nothing,
                          dynamics::Ptr{Cvoid}=C_NULL, user::Ptr{Cvoid}=C_NULL, plant=nothing, noise=nothing,
                          warm_start::Bool=true)
    mdl = g.mpc.model; obs = g.mpc.state_observer
    nx, nr, nd, nup, _ = LinearMPC.get_parameter_dims(g.mpc)
    kf = obs isa LinearMPC.OffsetFreeObserver ? obs.estimator : obs
    nxo = isnothing(kf) ? nx : size(kf.F, 1)
    ndm = nd - (nxo - nx)
    M = size(X0, 2)
    hold(v, k) = (k == 0 || isnothing(v)) ? nothing :
        (V = v isa AbstractMatrix ? Matrix{Float64}(v) : reshape(Vector{Float64}(v), k, 1);
         hcat(V[:, 1:min(N, size(V, 2))], repeat(V[:, end], 1, max(0, N - size(V, 2)))))
    Rt, Dt = hold(r, nr), hold(d, ndm)
    ny = isnothing(kf) ? 0 : size(kf.C, 1)
    Vt = hold(noise, ny)
    mats = isnothing(plant) ? (nothing, nothing, nothing, nothing) : plant
    Ft, Gt, Gdt, ft = (isnothing(m) ? nothing : Array{Float64}(m) for m in mats)
    p(x) = (isnothing(x) || isempty(x)) ? Ptr{Float64}(C_NULL) : pointer(x)
    Fo, Go, Gdo, fo, Co, Ddo, ho, Ko = isnothing(kf) ? ntuple(_ -> nothing, 8) :
        (Matrix{Float64}(kf.F), Matrix{Float64}(kf.G), Matrix{Float64}(kf.Gd), Vector{Float64}(kf.f_offset),
         Matrix{Float64}(kf.C), Matrix{Float64}(kf.Dd), Vector{Float64}(kf.h_offset), Matrix{Float64}(kf.K))
    Cm = isnothing(kf) ? nothing : Matrix{Float64}(mdl.C)
    Ddm = (isnothing(kf) || ndm == 0) ? nothing : Matrix{Float64}(mdl.Dd[:, 1:ndm])
    hm = isnothing(kf) ? nothing : Vector{Float64}(mdl.h_offset)
    Gdm = ndm == 0 ? nothing : Matrix{Float64}(mdl.Gd[:, 1:ndm])
    Xend = similar(X0); Xhat = Matrix{Float64}(undef, nxo, M)
    Ulast = Matrix{Float64}(undef, g.nu, M); flag = Vector{Int32}(undef, M)
    GC.@preserve Rt Dt Vt Ft Gt Gdt ft Fo Go Gdo fo Co Ddo ho Ko Cm Ddm hm Gdm Xhat begin
        cfg = Ref(LmpcSimStepped(dynamics, user, p(Ft), p(Gt), p(Gdt), p(ft), ny, p(Cm), p(Ddm), p(hm), p(Vt),
                                 nxo, ndm, p(Fo), p(Go), p(Gdo), p(fo), p(Co), p(Ddo), p(ho), p(Ko),
                                 p(Rt), p(Dt), pointer(Xhat)))
        check(ccall((:lmpc_gpu_simulate_batch_stepped, liblmpcgpu), Cint,
                    (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32, Int32, Int32,
                     Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{LmpcSimStepped},
                     Int64, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}),
                    g.handle, mdl.F, mdl.G, p(Gdm), nx, nr, nup,
                    X0, C_NULL, C_NULL, C_NULL, cfg, M, N, warm_start, Xend, Ulast, C_NULL, flag, C_NULL))
    end
    return Xend, Xhat, Ulast, flag
end

# ---- explicit controllers ---------------------------------------------------------------------
mutable struct GpuExplicitMPC
    handle::Ptr{Cvoid}
    empc::LinearMPC.ExplicitMPC
    nth::Int; nu::Int
end

"""
    setup_gpu!(empc; device=0, le_child=1, verify=256) -> GpuExplicitMPC

Uploads bst.halfplanes / feedbacks / hp_list / jump_list / clipping (needs `build_tree!(empc)` first,
src/explicit.jl:65-71).  The library's node arrays are 0-based with jump 0 marking a leaf; ParametricDAQP's
are 1-based with `next == id` marking a leaf, so hp_list is shifted here.

Branch orientation.  The library walks to `next` when a'θ ≤ b; ParametricDAQP is believed to walk to
`next + 1` (its generated C reads `id = next_id + (val <= b)`), which is `le_child = 1`.  ParametricDAQP
is not vendored in LinearMPC.jl, so instead of trusting that, the upload is CHECKED: `verify` parameters
sampled in the solution's box are evaluated by `ParametricDAQP.evaluate(bst, θ)` on the CPU and on the GPU;
if they disagree the other orientation is tried, and if neither agrees the call errors.
"""
function setup_gpu!(empc::LinearMPC.ExplicitMPC; device::Integer=0, le_child::Integer=1, verify::Integer=256)
    bst = empc.bst
    isnothing(bst) && error("Need to build a binary search tree to evaluate control law")  # src/utils.jl:54-55
    nth = size(bst.halfplanes, 1) - 1
    nu = size(bst.feedbacks[1], 2)
    fb = reduce(hcat, [vec(F) for F in bst.feedbacks])            # (nth+1)*nu × n_fb
    jump = Int32.(bst.jump_list)                                   # 0 stays 0 (leaf)
    hp = Int32.(bst.hp_list) .- Int32(1)
    clip = isnothing(bst.clipping) ? C_NULL : Matrix{Float64}(bst.clipping[:, 1:2])
    function upload(orient)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:lmpc_gpu_create_bst_oriented, liblmpcgpu), Cint,
                    (Ptr{Float64}, Int32, Ptr{Float64}, Int32, Ptr{Int32}, Ptr{Int32}, Int32, Ptr{Float64},
                     Int32, Int32, Int32, Cint, Ref{Ptr{Cvoid}}),
                    bst.halfplanes, size(bst.halfplanes, 2), fb, size(fb, 2), hp, jump, length(jump), clip,
                    nth, nu, orient, device, h))
        g = GpuExplicitMPC(h[], empc, nth, nu)
        finalizer(x -> ccall((:lmpc_gpu_destroy_bst, liblmpcgpu), Cint, (Ptr{Cvoid},), x.handle), g)
        return g
    end
    verify <= 0 && return upload(le_child)
    lb, ub = empc.TH.lb, empc.TH.ub                                # the box the solution was computed on
    Θ = lb .+ (ub .- lb) .* rand(nth, verify)
    clipu(u) = isnothing(bst.clipping) ? u : clamp.(u, bst.clipping[:, 1], bst.clipping[:, 2])
    Uref = reduce(hcat, [clipu(ParametricDAQP.evaluate(bst, Θ[:, k])) for k in 1:verify])
    for orient in (le_child, 1 - le_child)
        g = upload(orient)
        U = Matrix{Float64}(undef, nu, verify)
        check(ccall((:lmpc_gpu_evaluate_batch, liblmpcgpu), Cint,
                    (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Int32}), g.handle, Θ, verify, U, C_NULL))
        maximum(abs.(U .- Uref)) <= 1e-9 * max(1.0, maximum(abs.(Uref))) && return g
        orient == le_child && @warn "liblmpcgpu: tree orientation le_child=$le_child disagrees with ParametricDAQP.evaluate; trying the other one"
    end
    error("liblmpcgpu: the uploaded tree does not reproduce ParametricDAQP.evaluate in either branch orientation")
end

function LinearMPC.compute_control(g::GpuExplicitMPC, X::AbstractMatrix; r=nothing, d=nothing, uprev=nothing)
    nx, nr, nd, nup, _ = LinearMPC.get_parameter_dims(g.empc)
    N = size(X, 2)
    col(v, k) = v isa AbstractMatrix ? v[1:k, :] : repeat(isnothing(v) ? zeros(k) : v[1:k], 1, N)
    Θ = Matrix{Float64}([X; col(r, nr); col(d, nd); col(isnothing(uprev) ? g.empc.uprev : uprev, nup)])
    U = Matrix{Float64}(undef, g.nu, N)
    check(ccall((:lmpc_gpu_evaluate_batch, liblmpcgpu), Cint,
                (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Int32}), g.handle, Θ, N, U, C_NULL))
    g.empc.uprev .= U[:, end]
    return U
end

"""
    simulate(ge, X0; N, r, d) -> (Xend, Ulast)

`Simulation(empc, scenario)` (src/simulation.jl:37-116) for every column of X0, on the GPU: the tree
walk and the model's linear dynamics run on chip for N steps.
"""
function simulate(ge::GpuExplicitMPC, X0::Matrix{Float64}; N::Integer=100, r=nothing, d=nothing)
    mdl = ge.empc.model
    nx, nr, nd, nup, _ = LinearMPC.get_parameter_dims(ge.empc)
    M = size(X0, 2)
    traj(v, k) = (k == 0 || isnothing(v)) ? C_NULL :
        (V = v isa AbstractMatrix ? Matrix{Float64}(v) : reshape(Vector{Float64}(v), k, 1);
         hcat(V[:, 1:min(N, size(V, 2))], repeat(V[:, end], 1, max(0, N - size(V, 2)))))
    Rt, Dt = traj(r, nr), traj(d, nd)
    Xend = similar(X0); Ulast = Matrix{Float64}(undef, ge.nu, M)
    check(ccall((:lmpc_gpu_simulate_bst_batch, liblmpcgpu), Cint,
                (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32, Int32, Int32, Int32,
                 Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                 Int64, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                ge.handle, mdl.F, mdl.G, nd == 0 ? C_NULL : mdl.Gd, nx, nr, nd, nup,
                X0, C_NULL, C_NULL, C_NULL, Rt, Dt, M, N, Xend, Ulast, C_NULL))
    return Xend, Ulast
end

end # module
