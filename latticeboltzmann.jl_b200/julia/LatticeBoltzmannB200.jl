# LatticeBoltzmannB200.jl -- `ccall` binding that routes LatticeBoltzmann.jl's hot path
# (collide! -> stream! -> apply! and the moment evaluation of next!) to liblbm_b200.so (C ABI: include/lbm_b200.h).
#
# NOT EXECUTED in the build image (Julia is not installed there).  What keeps it honest: tests/test_abi.py checks every
# ccall in this file against the header (symbol, argument count AND argument types) and the struct layouts against the
# ctypes mirror; tests/c_abi_smoke.c drives the same entry points from C; the separable decomposition used by
# `cross_decompose` below is the algorithm of lbm/separable.py, which the CPU test-suite checks against every shipped
# problem's analytic fields.  The Python mirror in ../lbm/ binds the very same ABI and is what the GPU tests run.
#
# Usage (drop-in: the package's own entry points keep their signatures):
#   using LatticeBoltzmann, LatticeBoltzmannB200
#   LatticeBoltzmannB200.enable!()                    # simulate(problem, q; ...) now builds a B200Model
#   result = simulate(problem, D2Q9(); t_end = 1.0, collision_model = TRT)
#   result.processing_method.df[end]; result.f_stream
# or explicitly:
#   model = B200Model(problem, D2Q9(); collision_model = TRT)    # instead of LatticeBoltzmannModel
#   simulate(model, 0:n_steps)                                    # same call, same semantics
module LatticeBoltzmannB200

using LatticeBoltzmann
import LatticeBoltzmann: collide!, stream!, apply!, apply_boundary_conditions!, next!, should_stop!, simulate,
    CollisionModel, SRT, TRT, MRT, BounceBack, MovingWall, North, East, South, West, Quadrature, FluidFlowProblem,
    boundary_conditions, initialize, InitializationStrategy, ProcessingMethod, TrackHydrodynamicErrors,
    MeanVelocityStoppingCriteria, VelocityConvergenceStoppingCriteria, NoStoppingCriteria, delta_t, lattice_force,
    lattice_viscosity, lattice_velocity, density, velocity, pressure, deviatoric_tensor, dimension

const LIB = get(ENV, "LBM_B200_LIB", joinpath(@__DIR__, "..", "liblbm_b200.so"))
const LBM_MAX_TAU, LBM_MAX_BCS = 16, 8
const SEPARABLE_WINDOW = 2048      # lattice steps per separable force table

struct LbmBc                       # lbm_bc
    kind::Int32; direction::Int32
    x0::Int32; x1::Int32; y0::Int32; y1::Int32
    u::NTuple{2, Float64}; rho::Float64; T::Float64
end
struct LbmDesc                     # lbm_desc
    abi_version::Int32; nx::Int32; ny::Int32
    lattice::Int32; dtype::Int32; collision::Int32; arith::Int32
    ntau::Int32; tau::NTuple{LBM_MAX_TAU, Float64}
    n_bcs::Int32; bcs::NTuple{LBM_MAX_BCS, LbmBc}
    device::Int32; rank::Int32; world::Int32
    nccl_id::NTuple{128, UInt8}
end
struct LbmSepField                 # lbm_sep_field
    c0::Float64
    a::NTuple{2, Float64}
    x::NTuple{2, Ptr{Float64}}
    y::NTuple{2, Ptr{Float64}}
end
struct LbmBatchStop                # lbm_batch_stop
    kind::Int32; check_every::Int32
    tolerance::Float64
end

check(rc) = rc == 0 ? nothing : error(unsafe_string(ccall((:lbm_last_error, LIB), Cstring, ())))

lattice_id(::D2Q4) = 0; lattice_id(::D2Q5) = 1; lattice_id(::D2Q9) = 2; lattice_id(::D2Q13) = 3
lattice_id(::D2Q17) = 4; lattice_id(::D2Q21) = 5; lattice_id(::D2Q37) = 6
dir_id(::North) = 0; dir_id(::East) = 1; dir_id(::South) = 2; dir_id(::West) = 3

to_bc(bc::BounceBack) = LbmBc(0, dir_id(bc.direction), first(bc.xs), last(bc.xs), first(bc.ys), last(bc.ys),
                              (0.0, 0.0), 1.0, 1.0)
to_bc(bc::MovingWall) = LbmBc(1, dir_id(bc.direction), first(bc.xs), last(bc.xs), first(bc.ys), last(bc.ys),
                              (Float64(bc.u[1]), Float64(bc.u[2])), bc.ρ, bc.T)

cm_code(::SRT) = 0; cm_code(::TRT) = 1; cm_code(::MRT) = 2
taus(cm::SRT) = [cm.τ]; taus(cm::TRT) = [cm.τ_symmetric, cm.τ_asymmetric]; taus(cm::MRT) = collect(cm.τs)

function make_desc(nx, ny, q, code, t, bcs; dtype = Float64, exact = true, device = 0)
    pad(v, n, z) = ntuple(i -> i <= length(v) ? v[i] : z, n)
    zero_bc = LbmBc(0, 0, 0, 0, 0, 0, (0.0, 0.0), 1.0, 1.0)
    LbmDesc(1, nx, ny, lattice_id(q), dtype == Float64 ? 0 : 1, code, exact ? 0 : 1, length(t),
            pad(Float64.(t), LBM_MAX_TAU, 0.0), length(bcs), pad(map(to_bc, bcs), LBM_MAX_BCS, zero_bc), device, 0, 1,
            ntuple(_ -> 0x00, 128))
end

function create_context(desc::LbmDesc)
    ctx = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:lbm_create, LIB), Cint, (Ref{LbmDesc}, Ref{Ptr{Cvoid}}), Ref(desc), ctx))
    ctx[]
end

# ------------------------------------------------------------------------------------------------------------------
# the model: LatticeBoltzmannModel(problem, q; ...) (src/lattice_boltzmann_model.jl:15-33) with device-resident f
# ------------------------------------------------------------------------------------------------------------------
mutable struct B200Model{Q, CM, PM, BCs, P}
    ctx::Ptr{Cvoid}
    quadrature::Q
    collision_model::CM
    boundary_conditions::BCs
    processing_method::PM
    problem::P
    nx::Int; ny::Int
    force_window::UnitRange{Int}   # steps covered by the separable force table on the device
    force_static::Bool             # a time-independent force has been uploaded
end

function B200Model(problem, q; collision_model = SRT,
                   initialization_strategy = InitializationStrategy(problem), process_method = nothing,
                   dtype = Float64, exact = true, device = 0)
    cm = CollisionModel(collision_model, q, problem)
    bcs = boundary_conditions(problem)
    ctx = create_context(make_desc(problem.NX, problem.NY, q, cm_code(cm), taus(cm), bcs; dtype = dtype, exact = exact, device = device))
    f = initialize(initialization_strategy, q, problem, collision_model)   # Array{Float64,3}(NX, NY, Q)
    check(ccall((:lbm_upload_f, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx, f))
    model = B200Model(ctx, q, cm, bcs, process_method, problem, problem.NX, problem.NY, 0:-1, false)
    finalizer(m -> ccall((:lbm_destroy, LIB), Cvoid, (Ptr{Cvoid},), getfield(m, :ctx)), model)
    model
end

function f_stream(m::B200Model)
    f = Array{Float64}(undef, m.nx, m.ny, length(m.quadrature.weights))
    check(ccall((:lbm_download_f, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), m.ctx, f))
    f
end
function f_collision(m::B200Model)
    f = Array{Float64}(undef, m.nx, m.ny, length(m.quadrature.weights))
    check(ccall((:lbm_download_f_collision, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), m.ctx, f))
    f
end
Base.getproperty(m::B200Model, s::Symbol) =
    s === :f_stream ? f_stream(m) : s === :f_collision ? f_collision(m) : getfield(m, s)

# ------------------------------------------------------------------------------------------------------------------
# force: the closure `collision_model.force(x_idx, y_idx, time)` (srt.jl:52, trt.jl:77, mrt.jl:92) becomes data.
# Nothing about the problem is assumed: the closure is sampled on the grid at the first and last step of a batch and
# classified -- time-independent (uniform or per-node field), separable and time-dependent (F_x a function of (y, t), F_y
# of (x, t): DecayingShearFlow(static = true), decaying_shear_flow.jl:131-147; one table row per step), or neither (the
# batch is cut to single steps and the field is uploaded per step).
# ------------------------------------------------------------------------------------------------------------------
sample_force(m::B200Model, time) = [Float64(m.collision_model.force(x, y, time)[d]) for x in 1:m.nx, y in 1:m.ny, d in 1:2]
is_separable(F) = all(F[:, :, 1] .== F[1:1, :, 1]) && all(F[:, :, 2] .== F[:, 1:1, 2])

# Upload force data valid for steps t0 .. t0 + n - 1 and return how many of them (<= n) one lbm_step may take.
function prepare_force!(m::B200Model, t0::Int, n::Int, Δt)
    force = m.collision_model.force
    if force === nothing
        m.force_static || check(ccall((:lbm_set_force_none, LIB), Cint, (Ptr{Cvoid},), m.ctx))
        m.force_static = true
        return n
    end
    m.force_static && return n
    (t0 in m.force_window) && return min(n, last(m.force_window) - t0 + 1)
    F0 = sample_force(m, t0 * Δt)
    F1 = sample_force(m, (t0 + max(n, 2) - 1) * Δt)
    if F0 == F1 && F0 == sample_force(m, (t0 + 1) * Δt)        # time-independent
        if all(F0[:, :, 1] .== F0[1, 1, 1]) && all(F0[:, :, 2] .== F0[1, 1, 2])
            check(ccall((:lbm_set_force_uniform, LIB), Cint, (Ptr{Cvoid}, Cdouble, Cdouble), m.ctx, F0[1, 1, 1], F0[1, 1, 2]))
        else
            check(ccall((:lbm_set_force_field, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), m.ctx, F0))
        end
        m.force_static = true
        return n
    end
    if is_separable(F0) && is_separable(F1)                      # one (NY + NX)-entry row per step
        w = min(n, SEPARABLE_WINDOW)
        fx_of_y = [Float64(force(1, y, (t0 + k) * Δt)[1]) for y in 1:m.ny, k in 0:(w - 1)]   # column-major == C [k][y]
        fy_of_x = [Float64(force(x, 1, (t0 + k) * Δt)[2]) for x in 1:m.nx, k in 0:(w - 1)]   #              == C [k][x]
        check(ccall((:lbm_set_force_separable, LIB), Cint, (Ptr{Cvoid}, Int64, Int32, Ptr{Float64}, Ptr{Float64}),
                    m.ctx, t0, w, fx_of_y, fy_of_x))
        m.force_window = t0:(t0 + w - 1)
        return w
    end
    check(ccall((:lbm_set_force_field, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), m.ctx, F0))  # general closure: step by step
    return 1
end

function step!(m::B200Model, t0::Int, n::Int, Δt)
    done = 0
    while done < n
        k = prepare_force!(m, t0 + done, n - done, Δt)
        check(ccall((:lbm_step, LIB), Cint, (Ptr{Cvoid}, Int64, Int64, Cdouble), m.ctx, t0 + done, k, Δt))
        done += k
    end
end

# the generic functions of the hot loop on a model (src/lattice_boltzmann_model.jl:84-110)
function collide!(m::B200Model; time = 0.0)
    force = m.collision_model.force
    if force !== nothing      # one collide at an arbitrary `time`: a one-row separable table or the sampled field
        F = sample_force(m, time)
        check(ccall((:lbm_set_force_field, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), m.ctx, F))
        m.force_static = false; m.force_window = 0:-1
    else
        prepare_force!(m, 0, 1, 0.0)
    end
    check(ccall((:lbm_collide, LIB), Cint, (Ptr{Cvoid}, Int64, Cdouble), m.ctx, 0, time))
end
stream!(m::B200Model) = check(ccall((:lbm_stream, LIB), Cint, (Ptr{Cvoid},), m.ctx))
apply_boundary_conditions!(m::B200Model; time = 0.0) =
    check(ccall((:lbm_apply_bcs, LIB), Cint, (Ptr{Cvoid}, Cdouble), m.ctx, time))

# ------------------------------------------------------------------------------------------------------------------
# next!: processing methods on device-resident populations
# ------------------------------------------------------------------------------------------------------------------
next!(m::B200Model, t::Int64) = m.processing_method === nothing ? false : next!(m.processing_method, m, t)
# any processing method without a device version: hand it f_stream, as the reference does (:108-110)
next!(pm::ProcessingMethod, m::B200Model, t::Int64) = next!(pm, m.quadrature, f_stream(m), t)

# Does next!(pm, ..., t) read f or have side effects?  (Steps between two such t are one fused device batch.)
host_visible(pm, t) = true
host_visible(::Nothing, t) = false
host_visible(pm::TrackHydrodynamicErrors, t) =
    (mod(t, 100) == 0 && !(pm.stop_criteria isa NoStoppingCriteria)) || t == pm.n_steps || pm.should_process

function device_reduce(m::B200Model, kind)
    out = zeros(4)
    check(ccall((:lbm_reduce, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Int32), m.ctx, kind, out, 4))
    out
end
should_stop!(sc, m::B200Model) = should_stop!(sc, m.quadrature, f_stream(m))   # criteria without a device version
should_stop!(::NoStoppingCriteria, m::B200Model) = false
function should_stop!(sc::MeanVelocityStoppingCriteria, m::B200Model)           # stopping_criteria.jl:17-55
    s = device_reduce(m, 0)                       # sum u_x, node count
    u_mean = s[1] / s[2]
    converged = abs(u_mean / sc.old_mean_velocity - 1)
    converged < sc.tolerance && return true
    isnan(u_mean) && return true
    sc.old_mean_velocity = u_mean
    false
end
function should_stop!(sc::VelocityConvergenceStoppingCriteria, m::B200Model)    # stopping_criteria.jl:71-115
    s = device_reduce(m, 1)                       # sum |u - u_old|^2, sum |u_old|^2; u_old := u on the device
    converged = sqrt(s[1]) / s[2]
    (converged < sc.tolerance || isnan(converged))
end

# Skeleton (cross) decomposition of a field sampled on the grid into <= 2 products X(x) Y(y): exact up to round-off for
# every field of rank <= 2, which covers all analytic solutions the package ships.  Returns nothing when the residual
# is not at round-off level.  (Algorithm == lbm/separable.py `cross_decompose`, tested there.)
function cross_decompose(E::Matrix{Float64})
    R = copy(E)
    scale = maximum(abs, E)
    terms = Tuple{Float64, Vector{Float64}, Vector{Float64}}[]
    for _ in 1:2
        v, idx = findmax(abs.(R))
        v <= 1e-14 * max(scale, 1e-300) && break
        i, j = Tuple(idx)
        X, Y, pivot = R[:, j], R[i, :], R[i, j]
        push!(terms, (1 / pivot, X, Y))
        R .-= (X * Y') ./ pivot
    end
    maximum(abs, R) <= 1e-13 * max(scale, 1e-300) ? terms : nothing
end

# TrackHydrodynamicErrors.next! (track_hydrodynamic_errors.jl:52-221): the 16 sums come from the device
# (lbm_reduce_errors); the analytic fields are evaluated with the package's own functions and passed in separable form.
function next!(pm::TrackHydrodynamicErrors, m::B200Model, t::Int64)
    should_stop = mod(t, 100) == 0 && should_stop!(pm.stop_criteria, m)
    (!should_stop && t != pm.n_steps && !pm.should_process) && return false
    problem, q = pm.problem, m.quadrature
    nx, ny = m.nx, m.ny
    xr, yr = range(problem)
    time = t * delta_t(problem)
    Δ_ = nx == 1 ? (ny == 1 ? 1.0 : Float64(yr.step)) : (ny == 1 ? Float64(xr.step) : Float64(yr.step) * Float64(xr.step))
    fields = Matrix{Float64}[zeros(nx, ny) for _ in 1:8]   # rho, ux, uy, p, sxx, sxy, syx, syy
    for xi in 1:nx, yi in 1:ny
        x, y = xr[xi], yr[yi]
        u = velocity(problem, x, y, time)
        σ = deviatoric_tensor(q, problem, x, y, time)
        fields[1][xi, yi] = density(q, problem, x, y, time)
        fields[2][xi, yi] = u[1]; fields[3][xi, yi] = u[2]
        fields[4][xi, yi] = pressure(q, problem, x, y, time)
        fields[5][xi, yi] = σ[1, 1]; fields[6][xi, yi] = σ[1, 2]; fields[7][xi, yi] = σ[2, 1]; fields[8][xi, yi] = σ[2, 2]
    end
    decs = map(cross_decompose, fields)
    any(isnothing, decs) && return next!(pm, q, f_stream(m), t) | should_stop   # not separable: the reference's host path
    keep = Vector{Float64}[]
    sep = map(decs) do terms
        a = [0.0, 0.0]; px = [Ptr{Float64}(C_NULL), Ptr{Float64}(C_NULL)]; py = copy(px)
        for (k, (c, X, Y)) in enumerate(terms)
            push!(keep, X, Y)
            a[k] = c; px[k] = pointer(X); py[k] = pointer(Y)
        end
        LbmSepField(0.0, (a[1], a[2]), (px[1], px[2]), (py[1], py[2]))
    end
    s = zeros(16)
    τ = q.speed_of_sound_squared * lattice_viscosity(problem)
    GC.@preserve keep check(ccall((:lbm_reduce_errors, LIB), Cint, (Ptr{Cvoid}, Cdouble, Cdouble, Ptr{LbmSepField}, Ptr{Float64}),
                                  m.ctx, τ, problem.u_max, sep, s))
    push!(pm.df, (timestep = t, error_ρ = sqrt(s[1]), error_u = sqrt(s[2] / s[3]), error_p = sqrt(s[4] / s[5]),
                  error_σ_xx = sqrt(s[6] / s[7]), error_σ_xy = sqrt(s[8] / s[9]), error_σ_yy = sqrt(s[10] / s[11]),
                  error_σ_yx = sqrt(s[12] / s[13]), mass = Δ_ * s[14], momentum = Δ_ * s[15], energy = Δ_ * s[16]))
    should_stop
end

# TakeSnapshots.next! (take_snapshots.jl:12-29): push!(snapshots, copy(f_in)) becomes an asynchronous device-to-host copy
# (lbm_snapshot_begin): the array is pushed right away and filled while the next steps run; the copy is waited for when
# the next snapshot is taken and at the end of simulate (finish_snapshots!).
host_visible(pm::TakeSnapshots, t) = pm.every_t isa Int ? mod(t, pm.every_t) == 0 : t in pm.every_t
finish_snapshots!(m::B200Model) = check(ccall((:lbm_snapshot_end, LIB), Cint, (Ptr{Cvoid},), m.ctx))
function next!(pm::TakeSnapshots, m::B200Model, t::Int64)
    host_visible(pm, t) || return false
    finish_snapshots!(m)
    f = Array{Float64}(undef, m.nx, m.ny, length(m.quadrature.weights))
    push!(pm.snapshots, f)          # the vector keeps the array alive while the copy is in flight
    push!(pm.timesteps, t)
    check(ccall((:lbm_snapshot_begin, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), m.ctx, f))
    false
end

# CompareWithAnalyticalSolution.next! / process! (processing_methods.jl:100-262): the 12 sums of process! come from one
# device reduction (lbm_reduce_process); rho, u, p of the problem are passed in separable form like the error norms.
host_visible(pm::CompareWithAnalyticalSolution, t) =
    (mod(t, 100) == 0 && !(pm.stop_criteria isa NoStoppingCriteria)) || t == pm.n_steps || pm.should_process
function sep_fields(fields::Vector{Matrix{Float64}})
    decs = map(cross_decompose, fields)
    any(isnothing, decs) && return nothing, nothing
    keep = Vector{Float64}[]
    sep = map(decs) do terms
        a = [0.0, 0.0]; px = [Ptr{Float64}(C_NULL), Ptr{Float64}(C_NULL)]; py = copy(px)
        for (k, (c, X, Y)) in enumerate(terms)
            push!(keep, X, Y)
            a[k] = c; px[k] = pointer(X); py[k] = pointer(Y)
        end
        LbmSepField(0.0, (a[1], a[2]), (px[1], px[2]), (py[1], py[2]))
    end
    sep, keep
end
function b200_process!(pm::CompareWithAnalyticalSolution, m::B200Model, time::Float64)
    problem, q = pm.problem, m.quadrature
    nx, ny = m.nx, m.ny
    xr, yr = range(problem)
    fields = Matrix{Float64}[zeros(nx, ny) for _ in 1:8]   # rho, ux, uy, p (the last four stay zero)
    for xi in 1:nx, yi in 1:ny
        x, y = xr[xi], yr[yi]
        u = velocity(problem, x, y, time)
        fields[1][xi, yi] = density(q, problem, x, y, time)
        fields[2][xi, yi] = u[1]; fields[3][xi, yi] = u[2]
        fields[4][xi, yi] = pressure(q, problem, x, y, time)
    end
    sep, keep = sep_fields(fields)
    sep === nothing && return process!(problem, q, f_stream(m), time, pm.df)   # not separable: the reference's host path
    s = zeros(16)
    GC.@preserve keep check(ccall((:lbm_reduce_process, LIB), Cint, (Ptr{Cvoid}, Cdouble, Ptr{LbmSepField}, Ptr{Float64}),
                                  m.ctx, problem.u_max, sep, s))
    opp = Float64(yr.step) * Float64(xr.step)
    push!(pm.df, (density = s[1], momentum = s[2], total_energy = s[3], kinetic_energy = s[4], internal_energy = s[5],
                  density_a = s[6], momentum_a = s[7], total_energy_a = s[8], kinetic_energy_a = s[9], internal_energy_a = s[10],
                  error_u = sqrt(opp * s[11]), error_p = sqrt(opp * s[12]),
                  error_σ_xx = 0.0, error_σ_xy = 0.0, error_σ_yy = 0.0, error_σ_yx = 0.0))
    false
end
function next!(pm::CompareWithAnalyticalSolution, m::B200Model, t::Int64)
    time = t * delta_t(pm.problem)
    if mod(t, 100) == 0 && should_stop!(pm.stop_criteria, m)
        b200_process!(pm, m, time)
        return true
    end
    (!pm.should_process && t != pm.n_steps) && return false
    b200_process!(pm, m, time)
    false
end

# simulate(model, time) (src/lattice_boltzmann_model.jl:60-77): the steps between two host-visible next! points are ONE
# fused device batch (lbm_step); populations stay on the device throughout.
function simulate(m::B200Model, time)
    pm = m.processing_method
    Δt = pm !== nothing && isdefined(pm, :problem) ? delta_t(pm.problem) : 0.0
    t0, n = first(time), 0
    for t in time
        n += 1
        host_visible(pm, t + 1) || continue
        step!(m, t0, n, Δt)
        t0, n = t + 1, 0
        next!(m, t + 1) && (finish_snapshots!(m); return m)
    end
    n > 0 && step!(m, t0, n, Δt)
    next!(m, last(time) + 1)
    finish_snapshots!(m)
    m
end

# ------------------------------------------------------------------------------------------------------------------
# array-level operators: collide!(cm, q, f_in, f_out), stream!(q, f, f_new), apply!(bcs, q, f_new, f_old)
# (src/collision_models.jl:19, stream.jl:18-19, boundary_conditions.jl:6-12) on scratch contexts kept per
# (lattice, model, relaxation times, boundary conditions, shape)
# ------------------------------------------------------------------------------------------------------------------
const SCRATCH = Dict{Any, Ptr{Cvoid}}()
function scratch(q, code, t, bcs, nx, ny)
    get!(SCRATCH, (lattice_id(q), code, Tuple(t), Tuple(map(to_bc, bcs)), nx, ny)) do
        create_context(make_desc(nx, ny, q, code, t, bcs))
    end
end
function release_scratch!()
    foreach(ctx -> ccall((:lbm_destroy, LIB), Cvoid, (Ptr{Cvoid},), ctx), values(SCRATCH))
    empty!(SCRATCH)
end

function b200_collide!(cm, q::Quadrature, f_in::Array{Float64, 3}, f_out::Array{Float64, 3}; time = 0.0)
    nx, ny, _ = size(f_in)
    ctx = scratch(q, cm_code(cm), taus(cm), (), nx, ny)
    check(ccall((:lbm_upload_f, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx, f_in))
    if cm.force === nothing
        check(ccall((:lbm_set_force_none, LIB), Cint, (Ptr{Cvoid},), ctx))
    else
        F = [Float64(cm.force(x, y, time)[d]) for x in 1:nx, y in 1:ny, d in 1:2]
        check(ccall((:lbm_set_force_field, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx, F))
    end
    check(ccall((:lbm_collide, LIB), Cint, (Ptr{Cvoid}, Int64, Cdouble), ctx, 0, time))
    check(ccall((:lbm_download_f_collision, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx, f_out))
    f_out
end
function b200_stream!(q::Quadrature, f::Array{Float64, 3}, f_new::Array{Float64, 3})
    nx, ny, _ = size(f)
    ctx = scratch(q, 0, [1.0], (), nx, ny)
    check(ccall((:lbm_upload_f, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx, f))
    check(ccall((:lbm_upload_f_collision, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx, f))
    check(ccall((:lbm_stream, LIB), Cint, (Ptr{Cvoid},), ctx))
    check(ccall((:lbm_download_f, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx, f_new))
    f_new
end
function b200_apply!(bcs::AbstractVector, q::Quadrature, f_new::Array{Float64, 3}, f_old::Array{Float64, 3}; time = 0.0)
    nx, ny, _ = size(f_new)
    ctx = scratch(q, 0, [1.0], bcs, nx, ny)
    check(ccall((:lbm_upload_f, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx, f_new))
    check(ccall((:lbm_upload_f_collision, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx, f_old))
    check(ccall((:lbm_apply_bcs, LIB), Cint, (Ptr{Cvoid}, Cdouble), ctx, time))
    check(ccall((:lbm_download_f, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx, f_new))
    f_new
end

# ------------------------------------------------------------------------------------------------------------------
# page-locked arrays and asynchronous copies: a stream of independent jobs with several models in flight runs at the
# device rate (job k + 1's upload and job k - 1's download overlap job k's steps; bench.py's e2e figure)
# ------------------------------------------------------------------------------------------------------------------
# Array{Float64, 3}(nx, ny, Q) in page-locked host memory (lbm_host_alloc); freed by its finalizer
function pinned_array(nx::Integer, ny::Integer, nq::Integer)
    ptr = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:lbm_host_alloc, LIB), Cint, (Ref{Ptr{Cvoid}}, Csize_t), ptr, nx * ny * nq * sizeof(Float64)))
    f = unsafe_wrap(Array, Ptr{Float64}(ptr[]), (Int(nx), Int(ny), Int(nq)); own = false)
    finalizer(_ -> ccall((:lbm_host_free, LIB), Cint, (Ptr{Cvoid},), ptr[]), f)
    f
end
# f must be a pinned_array and stay alive (and, for uploads, unmodified) until sync!(m)
upload_async!(m::B200Model, f::Array{Float64, 3}) = check(ccall((:lbm_upload_f_async, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), m.ctx, f))
download_async!(m::B200Model, f::Array{Float64, 3}) = check(ccall((:lbm_download_f_async, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), m.ctx, f))
sync!(m::B200Model) = check(ccall((:lbm_sync, LIB), Cint, (Ptr{Cvoid},), m.ctx))
set_option!(m::B200Model, key::AbstractString, value::Integer) =
    check(ccall((:lbm_set_option, LIB), Cint, (Ptr{Cvoid}, Cstring, Int64), m.ctx, key, value))

# ------------------------------------------------------------------------------------------------------------------
# enable!(): the package's own entry points build / use the B200 path.  Done at run time (method overwriting is not
# allowed during precompilation): `simulate(problem, q; ...)` (lattice_boltzmann_model.jl:34-59) constructs a B200Model,
# and -- with array_ops = true -- the array-level generic functions run on the device as well.
# ------------------------------------------------------------------------------------------------------------------
function enable!(; array_ops = false, dtype = Float64, exact = true)
    @eval LatticeBoltzmann function simulate(problem::FluidFlowProblem, q::Quadrature; process_method = nothing,
                                             should_process = true,
                                             initialization_strategy = InitializationStrategy(problem), t_end = 1.0,
                                             collision_model = SRT)
        Δt = delta_t(problem)
        n_steps = round(Int, t_end / Δt)
        if isnothing(process_method)
            process_method = ProcessingMethod(problem, should_process, n_steps)
        end
        model = $B200Model(problem, q; collision_model = collision_model, initialization_strategy = initialization_strategy,
                           process_method = process_method, dtype = $dtype, exact = $exact)
        simulate(model, 0:n_steps)
    end
    if array_ops
        for CM in (SRT, TRT, MRT)
            @eval LatticeBoltzmann collide!(cm::$CM, q::Quadrature, f_in::Array{Float64, 3}, f_out::Array{Float64, 3}; time = 0.0) =
                $b200_collide!(cm, q, f_in, f_out; time = time)
        end
        @eval LatticeBoltzmann stream!(q::Quadrature, f::Array{Float64, 3}, f_new::Array{Float64, 3}) = $b200_stream!(q, f, f_new)
        @eval LatticeBoltzmann apply!(bcs::Vector{<:BoundaryCondition}, q::Quadrature, f_new::Array{Float64, 3},
                                      f_old::Array{Float64, 3}; time = 0.0) = $b200_apply!(bcs, q, f_new, f_old; time = time)
    end
    nothing
end

# ------------------------------------------------------------------------------------------------------------------
# batched sweeps: the loop of examples/notebooks/trt_magic_parameter.ipynb:30-103 as one device batch (lbm_batch_*)
# ------------------------------------------------------------------------------------------------------------------
# problems[k], τs[k] (a vector of relaxation times per solve); all problems of one shape.  Returns (timestep, sums) with
# sums[:, k] the 16 sums of TrackHydrodynamicErrors for solve k at its stop step.
function simulate_many(problems::Vector, q::Quadrature, τs::Vector{<:AbstractVector}; collision_model = TRT, t_end = 1.0,
                       stop_kind = 2, tolerance = 1e-7, check_every = 100, expected = nothing)
    p0 = problems[1]
    B, ntau = length(problems), length(τs[1])
    bcs = boundary_conditions(p0)
    code = collision_model === SRT ? 0 : collision_model === TRT ? 1 : 2
    batch = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:lbm_batch_create, LIB), Cint, (Ref{LbmDesc}, Int32, Ref{Ptr{Cvoid}}),
                Ref(make_desc(p0.NX, p0.NY, q, code, τs[1], bcs)), B, batch))
    try
        tau = [Float64(τs[k][i]) for i in 1:ntau, k in 1:B]                               # column-major == C [k][i]
        check(ccall((:lbm_batch_set_tau, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), batch[], tau))
        if LatticeBoltzmann.has_external_force(p0)
            F = [Float64(lattice_force(problems[k], 1, 1, 0.0)[d]) for d in 1:2, k in 1:B]
            check(ccall((:lbm_batch_set_force_uniform, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), batch[], F))
        end
        f0 = initialize(LatticeBoltzmann.ZeroVelocityInitialCondition(), q, p0)
        check(ccall((:lbm_batch_broadcast_f, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), batch[], f0))
        n_steps = round(Int, t_end / delta_t(p0))
        stop = Ref(LbmBatchStop(stop_kind, check_every, tolerance))
        check(ccall((:lbm_batch_run, LIB), Cint, (Ptr{Cvoid}, Int64, Ref{LbmBatchStop}), batch[], n_steps, stop))
        timestep, stopped = zeros(Int64, B), zeros(Int32, B)
        check(ccall((:lbm_batch_status, LIB), Cint, (Ptr{Cvoid}, Int32, Int32, Ptr{Int64}, Ptr{Int32}), batch[], 0, B, timestep, stopped))
        sums = zeros(16, B)
        if expected !== nothing   # (sep::Vector{LbmSepField} with shared tables, coef::Array{Float64,3} (3, 8, B), keep)
            sep, coef, keep = expected
            τv = [q.speed_of_sound_squared * lattice_viscosity(p) for p in problems]
            um = [Float64(p.u_max) for p in problems]
            GC.@preserve keep check(ccall((:lbm_batch_reduce_errors, LIB), Cint,
                                          (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{LbmSepField}, Ptr{Float64}, Ptr{Float64}),
                                          batch[], τv, um, sep, coef, sums))
        end
        return timestep, stopped, sums
    finally
        ccall((:lbm_batch_destroy, LIB), Cvoid, (Ptr{Cvoid},), batch[])
    end
end

# initialize(::IterativeInitializationMeiEtAl, q, problem) (src/initial_conditions/mei_et_al.jl:11-40) on the device:
# collision kind 3 = IterativeInitializationCollisionModel (src/collision_models/iterative_initialization.jl), the
# prescribed lattice velocity goes in once, DensityConvergence (stopping_criteria/density_convergence.jl:6-17) reads the
# density of node (NX, NY) -- the only node its loop visits -- from lbm_reduce kind 3 (out[2]).
function initialize_mei_et_al(strategy, q, problem)
    bcs = boundary_conditions(problem)
    ctx = create_context(make_desc(problem.NX, problem.NY, q, 3, [Float64(strategy.τ)], bcs))
    xs, ys = range(problem)
    u0 = [lattice_velocity(q, problem, xs[x], ys[y])[d] for x in 1:problem.NX, y in 1:problem.NY, d in 1:2]
    f = [q.weights[i] for x in 1:problem.NX, y in 1:problem.NY, i in 1:length(q.weights)]
    out, ρ_old = zeros(4), 0.0
    try
        check(ccall((:lbm_upload_f, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx, f))
        check(ccall((:lbm_set_velocity_field, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx, u0))
        for step in 1:10000
            check(ccall((:lbm_step, LIB), Cint, (Ptr{Cvoid}, Int64, Int64, Cdouble), ctx, step, 1, 0.0))
            check(ccall((:lbm_reduce, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Int32), ctx, 3, out, 4))
            δρ = abs(out[2] - ρ_old); ρ_old = out[2]
            (δρ < strategy.ϵ || δρ > 100.0) && break
        end
        check(ccall((:lbm_download_f, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx, f))
    finally
        ccall((:lbm_destroy, LIB), Cvoid, (Ptr{Cvoid},), ctx)
    end
    f
end

# 0 single GPU, 1 NCCL send/recv, 2 peer-memory stores issued by the boundary-row kernel
halo_path(m::B200Model) = ccall((:lbm_halo_path, LIB), Cint, (Ptr{Cvoid},), m.ctx)

export B200Model
end # module
