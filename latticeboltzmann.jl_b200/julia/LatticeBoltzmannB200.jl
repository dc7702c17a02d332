# LatticeBoltzmannB200.jl -- `ccall` shim that routes LatticeBoltzmann.jl's hot path
# (collide! -> stream! -> apply! and the moment evaluation of next!) to liblbm_b200.so.
#
# NOT EXECUTED in the build image (Julia is not installed there); it documents, line for line,
# the binding a maintainer adds.  The Python mirror in ../lbm/ binds the very same C ABI
# (include/lbm_b200.h) and is what the test-suite runs.
#
# Usage:
#   using LatticeBoltzmann, LatticeBoltzmannB200
#   model = B200Model(problem, D2Q9(); collision_model = TRT)   # instead of LatticeBoltzmannModel
#   simulate(model, 0:n_steps)                                   # same call, same semantics
module LatticeBoltzmannB200

using LatticeBoltzmann
import LatticeBoltzmann: collide!, stream!, apply_boundary_conditions!, next!, simulate,
    CollisionModel, SRT, TRT, MRT, BounceBack, MovingWall, North, East, South, West,
    boundary_conditions, initialize, InitializationStrategy, delta_t, lattice_force, order

const LIB = get(ENV, "LBM_B200_LIB", joinpath(@__DIR__, "..", "liblbm_b200.so"))
const LBM_MAX_TAU, LBM_MAX_BCS = 16, 8

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

mutable struct B200Model{Q, CM, PM, BCs}
    ctx::Ptr{Cvoid}
    quadrature::Q
    collision_model::CM
    boundary_conditions::BCs
    processing_method::PM
    nx::Int; ny::Int
end

function B200Model(problem, q; collision_model = SRT,
                   initialization_strategy = InitializationStrategy(problem), process_method = nothing,
                   dtype = Float64, exact = true)
    cm = CollisionModel(collision_model, q, problem)
    bcs = boundary_conditions(problem)
    t = taus(cm)
    pad(v, n, z) = ntuple(i -> i <= length(v) ? v[i] : z, n)
    zero_bc = LbmBc(0, 0, 0, 0, 0, 0, (0.0, 0.0), 1.0, 1.0)
    desc = Ref(LbmDesc(1, problem.NX, problem.NY, lattice_id(q), dtype == Float64 ? 0 : 1, cm_code(cm),
                       exact ? 0 : 1, length(t), pad(t, LBM_MAX_TAU, 0.0), length(bcs),
                       pad(map(to_bc, bcs), LBM_MAX_BCS, zero_bc), 0, 0, 1, ntuple(_ -> 0x00, 128)))
    ctx = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:lbm_create, LIB), Cint, (Ref{LbmDesc}, Ref{Ptr{Cvoid}}), desc, ctx))
    f = initialize(initialization_strategy, q, problem, collision_model)   # Array{Float64,3}(NX, NY, Q)
    check(ccall((:lbm_upload_f, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx[], f))
    model = B200Model(ctx[], q, cm, bcs, process_method, problem.NX, problem.NY)
    finalizer(m -> ccall((:lbm_destroy, LIB), Cvoid, (Ptr{Cvoid},), m.ctx), model)
    set_force!(model, problem)
    model
end

# The force closure becomes data (lbm_set_force_*): uniform for index-based forces (Poiseuille),
# a static field otherwise; DecayingShearFlow uses lbm_set_force_separable per batch.
function set_force!(m::B200Model, problem)
    m.collision_model.force === nothing && return check(ccall((:lbm_set_force_none, LIB), Cint, (Ptr{Cvoid},), m.ctx))
    F = [lattice_force(problem, x, y, 0.0)[d] for x in 1:m.nx, y in 1:m.ny, d in 1:2]
    if all(F[:, :, 1] .== F[1, 1, 1]) && all(F[:, :, 2] .== F[1, 1, 2])
        check(ccall((:lbm_set_force_uniform, LIB), Cint, (Ptr{Cvoid}, Cdouble, Cdouble), m.ctx, F[1, 1, 1], F[1, 1, 2]))
    else
        check(ccall((:lbm_set_force_field, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), m.ctx, F))
    end
end

f_stream(m::B200Model) = (f = Array{Float64}(undef, m.nx, m.ny, length(m.quadrature.weights));
    check(ccall((:lbm_download_f, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), m.ctx, f)); f)
Base.getproperty(m::B200Model, s::Symbol) = s === :f_stream ? f_stream(m) : getfield(m, s)

# the four generic functions of the hot loop (src/lattice_boltzmann_model.jl:84-110)
collide!(m::B200Model; time) = check(ccall((:lbm_collide, LIB), Cint, (Ptr{Cvoid}, Int64, Cdouble), m.ctx, 0, time))
stream!(m::B200Model) = check(ccall((:lbm_stream, LIB), Cint, (Ptr{Cvoid},), m.ctx))
apply_boundary_conditions!(m::B200Model; time = 0.0) =
    check(ccall((:lbm_apply_bcs, LIB), Cint, (Ptr{Cvoid}, Cdouble), m.ctx, time))
next!(m::B200Model, t::Int64) = m.processing_method === nothing ? false :
    next!(m.processing_method, m.quadrature, f_stream(m), t)   # or a device-side method using lbm_moments/lbm_reduce

# simulate(model, time) (src/lattice_boltzmann_model.jl:60-77) with the steps between two
# host-visible next! points issued as ONE fused device batch (lbm_step).
function simulate(m::B200Model, time)
    pm = m.processing_method
    Δt = pm !== nothing && isdefined(pm, :problem) ? delta_t(pm.problem) : 0.0
    host_visible(t) = pm !== nothing && (mod(t, 100) == 0 || t == pm.n_steps || pm.should_process)
    t0, n = first(time), 0
    for t in time
        n += 1
        host_visible(t + 1) || continue
        check(ccall((:lbm_step, LIB), Cint, (Ptr{Cvoid}, Int64, Int64, Cdouble), m.ctx, t0, n, Δt))
        t0, n = t + 1, 0
        next!(m, t + 1) && return m
    end
    n > 0 && check(ccall((:lbm_step, LIB), Cint, (Ptr{Cvoid}, Int64, Int64, Cdouble), m.ctx, t0, n, Δt))
    next!(m, last(time) + 1)
    m
end

# initialize(::IterativeInitializationMeiEtAl, q, problem) (src/initial_conditions/mei_et_al.jl:11-40) on the device:
# collision kind 3 = IterativeInitializationCollisionModel (src/collision_models/iterative_initialization.jl), the
# prescribed lattice velocity goes in once, DensityConvergence (stopping_criteria/density_convergence.jl:6-17) reads the
# density of node (NX, NY) -- the only node its loop visits -- from lbm_reduce kind 3 (out[2]).
function initialize_mei_et_al(strategy, q, problem)
    t = pad -> ntuple(i -> i == 1 ? Float64(strategy.τ) : 0.0, pad)
    zero_bc = LbmBc(0, 0, 0, 0, 0, 0, (0.0, 0.0), 1.0, 1.0)
    bcs = boundary_conditions(problem)
    desc = Ref(LbmDesc(1, problem.NX, problem.NY, lattice_id(q), 0, 3, 0, 1, t(LBM_MAX_TAU), length(bcs),
                       ntuple(i -> i <= length(bcs) ? to_bc(bcs[i]) : zero_bc, LBM_MAX_BCS), 0, 0, 1, ntuple(_ -> 0x00, 128)))
    ctx = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:lbm_create, LIB), Cint, (Ref{LbmDesc}, Ref{Ptr{Cvoid}}), desc, ctx))
    xs, ys = range(problem)
    u0 = [lattice_velocity(q, problem, xs[x], ys[y])[d] for x in 1:problem.NX, y in 1:problem.NY, d in 1:2]
    f = [q.weights[i] for x in 1:problem.NX, y in 1:problem.NY, i in 1:length(q.weights)]
    out, ρ_old = zeros(4), 0.0
    try
        check(ccall((:lbm_upload_f, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx[], f))
        check(ccall((:lbm_set_velocity_field, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx[], u0))
        for step in 1:10000
            check(ccall((:lbm_step, LIB), Cint, (Ptr{Cvoid}, Int64, Int64, Cdouble), ctx[], step, 1, 0.0))
            check(ccall((:lbm_reduce, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Int32), ctx[], 3, out, 4))
            δρ = abs(out[2] - ρ_old); ρ_old = out[2]
            (δρ < strategy.ϵ || δρ > 100.0) && break
        end
        check(ccall((:lbm_download_f, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx[], f))
    finally
        ccall((:lbm_destroy, LIB), Cvoid, (Ptr{Cvoid},), ctx[])
    end
    f
end

# 0 single GPU, 1 NCCL send/recv, 2 peer-memory stores issued by the boundary-row kernel
halo_path(m::B200Model) = ccall((:lbm_halo_path, LIB), Cint, (Ptr{Cvoid},), m.ctx)

export B200Model
end # module
