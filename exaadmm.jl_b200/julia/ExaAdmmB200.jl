# ExaAdmmB200.jl — Julia glue for the B200-native ACOPF path (UNEXECUTED: the build image
# has no Julia; the same C ABI is exercised from Python ctypes, see exaadmm.jl_b200/capi.py).
#
# How it plugs into ExaAdmm.jl: the reference selects a backend purely by the array type
# parameters of (AdmmEnv{T,TD,TI,TM}, AbstractOPFModel{T,TD,TI,TM}) (e.g.
# src/models/acopf/acopf_admm_update_x_gpu.jl:48-52). This file adds one more tag,
# `B200Vector`, and one method per operator whose body is a single `ccall` into
# libexaadmm_b200.so (include/exaadmm_b200.h). `solve_acopf(...; use_gpu=true)` keeps
# its signature and its `(env, mod)` return; INTEGRATION.md shows the 3-line patch to
# src/interface/solve_acopf.jl that routes `use_gpu=true` here.
module ExaAdmmB200

using ExaAdmm
# ExaAdmm exports only its solve_* entry points (src/ExaAdmm.jl), so every type and generic function that gets a method
# here is imported by name: a method defined on an un-imported name would create a new local function instead of
# extending ExaAdmm's.
import ExaAdmm: AdmmEnv, AbstractOPFModel, AbstractSolution, ModelAcopf, ModelMpacopf, ModelQpsub, GridData, Solution,
                IterationInformation, ComponentInformation, admm_increment_outer, admm_increment_reset_inner,
                admm_increment_inner, admm_outer_prestep, admm_inner_prestep, admm_update_x, admm_update_xbar,
                admm_update_z, admm_update_l, admm_update_l_single, admm_update_residual, admm_update_lz, admm_poststep,
                admm_two_level, admm_one_level, init_solution!, print_statistics

const LIB = get(ENV, "EXAADMM_B200_LIB", "libexaadmm_b200.so")

# ---- array-type tag: a host shadow + the library handle that owns the HBM copy ----------
mutable struct B200Vector{T} <: AbstractVector{T}
    handle::Ptr{Cvoid}      # ea_handle_t*, shared by all vectors of one model
    field::Cint             # enum ea_field
    n::Int
end
Base.size(v::B200Vector) = (v.n,)
function Base.copyto!(dst::Array{Float64,1}, src::B200Vector{Float64})           # copyto!(host, mod.solution.u_curr)
    rc = ccall((:ea_get_vector, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Int64), src.handle, src.field, dst, length(dst))
    rc == 0 || error(unsafe_string(ccall((:ea_last_error, LIB), Cstring, (Ptr{Cvoid},), src.handle)))
    return dst
end
function Base.copyto!(dst::B200Vector{Float64}, src::Array{Float64,1})
    rc = ccall((:ea_set_vector, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Int64), dst.handle, dst.field, src, length(src))
    rc == 0 || error(unsafe_string(ccall((:ea_last_error, LIB), Cstring, (Ptr{Cvoid},), dst.handle)))
    return dst
end
Base.getindex(v::B200Vector{Float64}, i::Int) = copyto!(zeros(v.n), v)[i]       # debugging convenience only

const TD = B200Vector{Float64}
const TI = Array{Int,1}          # index arrays stay on the host; the library copies them once in ea_create
const TM = Array{Float64,2}
const Env = AdmmEnv{Float64,TD,TI,TM}
const Mod = AbstractOPFModel{Float64,TD,TI,TM}

# ---- ea_grid_t (field order of include/exaadmm_b200.h) ------------------------------------
struct EaGrid
    ngen::Int64; nline::Int64; nbus::Int64; baseMVA::Cdouble
    pgmin::Ptr{Cdouble}; pgmax::Ptr{Cdouble}; qgmin::Ptr{Cdouble}; qgmax::Ptr{Cdouble}
    c2::Ptr{Cdouble}; c1::Ptr{Cdouble}; c0::Ptr{Cdouble}
    YshR::Ptr{Cdouble}; YshI::Ptr{Cdouble}
    YffR::Ptr{Cdouble}; YffI::Ptr{Cdouble}; YftR::Ptr{Cdouble}; YftI::Ptr{Cdouble}
    YttR::Ptr{Cdouble}; YttI::Ptr{Cdouble}; YtfR::Ptr{Cdouble}; YtfI::Ptr{Cdouble}
    FrVmBound::Ptr{Cdouble}; ToVmBound::Ptr{Cdouble}; FrVaBound::Ptr{Cdouble}; ToVaBound::Ptr{Cdouble}
    rateA::Ptr{Cdouble}
    FrStart::Ptr{Int64}; ToStart::Ptr{Int64}; GenStart::Ptr{Int64}
    FrIdx::Ptr{Int64}; ToIdx::Ptr{Int64}; GenIdx::Ptr{Int64}
    Pd::Ptr{Cdouble}; Qd::Ptr{Cdouble}; Vmin::Ptr{Cdouble}; Vmax::Ptr{Cdouble}
    brBusIdx::Ptr{Int64}
end

struct EaParams
    mu_max::Cdouble; max_auglag::Int32; verbose::Int32
    initial_beta::Cdouble; inc_c::Cdouble; theta::Cdouble; outer_eps::Cdouble
    MAX_MULTIPLIER::Cdouble; scale::Cdouble; obj_scale::Cdouble
    outer_iterlim::Int64; inner_iterlim::Int64
end

mutable struct EaInfo
    status::Int32; _pad::Int32
    inner::Int64; outer::Int64; cumul::Int64
    objval::Cdouble; primres::Cdouble; dualres::Cdouble; mismatch::Cdouble; auglag::Cdouble; eps_pri::Cdouble
    norm_z_curr::Cdouble; norm_z_prev::Cdouble; beta::Cdouble
    time_x_update::Cdouble; time_xbar_update::Cdouble; time_z_update::Cdouble; time_l_update::Cdouble
    time_lz_update::Cdouble; time_projection::Cdouble; time_overall::Cdouble
    time_generators::Cdouble; time_branches::Cdouble; time_buses::Cdouble
    EaInfo() = new()
end

check(h, rc) = rc == 0 || error("exaadmm_b200 ($rc): " * unsafe_string(ccall((:ea_last_error, LIB), Cstring, (Ptr{Cvoid},), h)))
handle(mod::Mod) = mod.solution.u_curr.handle

"""
    create_handle(g::GridData{Float64,Array{Float64,1},Array{Int,1},Array{Float64,2}}, gpu_no) -> Ptr{Cvoid}

The H2D part of the `ModelAcopf` constructor (acopf_model.jl:41-94): pass the host `GridData`
(built with the reference's own loader, use_gpu=false) to `ea_create`.
"""
# `ea_grid_t` over the host arrays of `g` (valid while `g` is alive: callers wrap the ccall in GC.@preserve g)
grid_struct(g) = EaGrid(g.ngen, g.nline, g.nbus, g.baseMVA,
    pointer(g.pgmin), pointer(g.pgmax), pointer(g.qgmin), pointer(g.qgmax), pointer(g.c2), pointer(g.c1), pointer(g.c0),
    pointer(g.YshR), pointer(g.YshI), pointer(g.YffR), pointer(g.YffI), pointer(g.YftR), pointer(g.YftI),
    pointer(g.YttR), pointer(g.YttI), pointer(g.YtfR), pointer(g.YtfI),
    pointer(g.FrVmBound), pointer(g.ToVmBound), pointer(g.FrVaBound), pointer(g.ToVaBound), pointer(g.rateA),
    pointer(g.FrStart), pointer(g.ToStart), pointer(g.GenStart), pointer(g.FrIdx), pointer(g.ToIdx), pointer(g.GenIdx),
    pointer(g.Pd), pointer(g.Qd), pointer(g.Vmin), pointer(g.Vmax), pointer(g.brBusIdx))

function create_handle(g, gpu_no::Int)
    GC.@preserve g begin
        out = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:ea_create, LIB), Cint, (Ref{EaGrid}, Cint, Ref{Ptr{Cvoid}}), grid_struct(g), gpu_no, out)
        rc == 0 || error("ea_create ($rc): " * unsafe_string(ccall((:ea_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
        return out[]
    end
end

# An owned library handle kept in a model field (`gen_solution` slot of the multi-period and QP models); the finalizer
# calls the matching destructor (:ea_destroy, :ea_mp_destroy or :ea_qp_destroy).
mutable struct B200Handle <: AbstractSolution{Float64,TD}
    handle::Ptr{Cvoid}
    function B200Handle(h::Ptr{Cvoid}, destructor::Symbol)
        obj = new(h)
        finalizer(obj) do o
            o.handle == C_NULL && return
            destructor === :ea_qp_destroy ? ccall((:ea_qp_destroy, LIB), Cvoid, (Ptr{Cvoid},), o.handle) :
            destructor === :ea_mp_destroy ? ccall((:ea_mp_destroy, LIB), Cvoid, (Ptr{Cvoid},), o.handle) :
                                            ccall((:ea_destroy, LIB), Cvoid, (Ptr{Cvoid},), o.handle)
            o.handle = C_NULL
        end
        return obj
    end
end

# ---- one ccall per operator (docs/src/dev.md:65-155) ---------------------------------------
function init_solution!(mod::Mod, sol::Solution{Float64,TD}, rho_pq::Float64, rho_va::Float64, device=nothing)
    check(handle(mod), ccall((:ea_init_solution, LIB), Cint, (Ptr{Cvoid}, Cdouble, Cdouble), handle(mod), rho_pq, rho_va))
end
function admm_outer_prestep(env::Env, mod::Mod, device=nothing)
    out = Ref{Cdouble}(0)
    check(handle(mod), ccall((:ea_outer_prestep, LIB), Cint, (Ptr{Cvoid}, Ref{Cdouble}), handle(mod), out))
    mod.info.norm_z_prev = out[]; return
end
admm_inner_prestep(env::Env, mod::Mod, device=nothing) =
    (check(handle(mod), ccall((:ea_inner_prestep, LIB), Cint, (Ptr{Cvoid},), handle(mod))); nothing)
function admm_update_x(env::Env, mod::Mod, device=nothing)
    p = env.params
    check(handle(mod), ccall((:ea_update_x, LIB), Cint, (Ptr{Cvoid}, Int64, Int32, Cdouble, Cdouble),
                             handle(mod), mod.info.inner, p.max_auglag, p.mu_max, p.scale)); return
end
admm_update_xbar(env::Env, mod::Mod, device=nothing) =
    (check(handle(mod), ccall((:ea_update_xbar, LIB), Cint, (Ptr{Cvoid},), handle(mod))); nothing)
admm_update_z(env::Env, mod::Mod, device=nothing) =
    (check(handle(mod), ccall((:ea_update_z, LIB), Cint, (Ptr{Cvoid}, Cdouble), handle(mod), env.params.beta)); nothing)
admm_update_l(env::Env, mod::Mod, device=nothing) =
    (check(handle(mod), ccall((:ea_update_l, LIB), Cint, (Ptr{Cvoid}, Cdouble), handle(mod), env.params.beta)); nothing)
admm_update_lz(env::Env, mod::Mod, device=nothing) =
    (check(handle(mod), ccall((:ea_update_lz, LIB), Cint, (Ptr{Cvoid}, Cdouble, Cdouble), handle(mod), env.params.beta, env.params.MAX_MULTIPLIER)); nothing)
function admm_update_residual(env::Env, mod::Mod, device=nothing)
    out = zeros(4)
    check(handle(mod), ccall((:ea_update_residual, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), handle(mod), out))
    mod.info.primres, mod.info.dualres, mod.info.norm_z_curr, mod.info.mismatch = out; return
end
function admm_poststep(env::Env, mod::Mod, device=nothing)
    out = Ref{Cdouble}(0)
    check(handle(mod), ccall((:ea_poststep, LIB), Cint, (Ptr{Cvoid}, Ref{Cdouble}), handle(mod), out))
    mod.info.objval = out[]; return
end

# ---- fast path: the whole two-level loop in one ccall ----------------------------------------
# admm_two_level is generic in the reference (src/algorithms/admm_two_level.jl:1-88); this method is
# more specific (array tag B200Vector), so dispatch picks it for the B200 model and the inner loop
# never returns to Julia.
function admm_two_level(env::Env, mod::Mod, device=nothing)
    p = env.params
    par = EaParams(p.mu_max, p.max_auglag, p.verbose, p.initial_beta, p.inc_c, p.theta, p.outer_eps,
                   p.MAX_MULTIPLIER, p.scale, p.obj_scale, p.outer_iterlim, p.inner_iterlim)
    info = EaInfo()
    check(handle(mod), ccall((:ea_admm_two_level, LIB), Cint, (Ptr{Cvoid}, Ref{EaParams}, Ref{EaInfo}), handle(mod), par, info))
    i = mod.info
    i.status = (:NotSpecified, :IterationLimit, :Solved)[info.status + 1]
    i.inner, i.outer, i.cumul = info.inner, info.outer, info.cumul
    i.objval, i.primres, i.dualres, i.mismatch = info.objval, info.primres, info.dualres, info.mismatch
    i.eps_pri, i.norm_z_curr, i.norm_z_prev = info.eps_pri, info.norm_z_curr, info.norm_z_prev
    i.time_x_update, i.time_xbar_update, i.time_z_update = info.time_x_update, info.time_xbar_update, info.time_z_update
    i.time_l_update, i.time_lz_update, i.time_overall = info.time_l_update, info.time_lz_update, info.time_overall
    i.user.time_generators, i.user.time_branches, i.user.time_buses = info.time_generators, info.time_branches, info.time_buses
    p.beta = info.beta
    p.verbose > 0 && print_statistics(env, mod)
    return
end

# ---- multi-period model: ModelMpacopf{Float64,TD,TI,TM} (src/models/mpacopf/) ---------------------
# The model's constructor calls ea_mp_create (loads period-major) and wraps ea_mp_period(h, t) as the handles of
# mod.models[t+1]; mod.solution[i].<field> are B200RampVector(handle, period, field) with copyto! ->
# ea_mp_get_ramp_vector / ea_mp_set_ramp_vector. One method per generic function, as for the single-period model:
const MpMod = ModelMpacopf{Float64,TD,TI,TM}
mp_handle(mod::MpMod) = mod.models[1].gen_solution.handle      # the ea_mp_handle_t* is kept in the gen_solution slot
mp_check(h, rc) = rc == 0 || error(unsafe_string(ccall((:ea_mp_last_error, LIB), Cstring, (Ptr{Cvoid},), h)))

# Without these four the calls would fall through to the single-period methods above, whose handle(mod) reads
# mod.solution.u_curr - a Vector of ramp solutions for this model.
function init_solution!(mod::MpMod, sol, rho_pq::Float64, rho_va::Float64, device=nothing)
    mp_check(mp_handle(mod), ccall((:ea_mp_init_solution, LIB), Cint, (Ptr{Cvoid}, Cdouble, Cdouble), mp_handle(mod), rho_pq, rho_va)); return
end
function admm_outer_prestep(env::Env, mod::MpMod, device=nothing)
    out = Ref{Cdouble}(0)
    mp_check(mp_handle(mod), ccall((:ea_mp_outer_prestep, LIB), Cint, (Ptr{Cvoid}, Ref{Cdouble}), mp_handle(mod), out))
    mod.info.norm_z_prev = out[]; return
end
admm_inner_prestep(env::Env, mod::MpMod, device=nothing) =
    (mp_check(mp_handle(mod), ccall((:ea_mp_inner_prestep, LIB), Cint, (Ptr{Cvoid},), mp_handle(mod))); nothing)
function admm_poststep(env::Env, mod::MpMod, device=nothing)
    obj = Ref{Cdouble}(0); err = Ref{Cdouble}(0)
    mp_check(mp_handle(mod), ccall((:ea_mp_poststep, LIB), Cint, (Ptr{Cvoid}, Ref{Cdouble}, Ref{Cdouble}), mp_handle(mod), obj, err))
    mod.info.objval = obj[]; mod.info.user.err_ramp = err[]; return
end
function admm_update_x(env::Env, mod::MpMod, device=nothing)
    p = env.params
    mp_check(mp_handle(mod), ccall((:ea_mp_update_x, LIB), Cint, (Ptr{Cvoid}, Int64, Int32, Cdouble, Cdouble),
                                   mp_handle(mod), mod.info.inner, p.max_auglag, p.mu_max, p.scale)); return
end
admm_update_xbar(env::Env, mod::MpMod, device=nothing) =
    (mp_check(mp_handle(mod), ccall((:ea_mp_update_xbar, LIB), Cint, (Ptr{Cvoid},), mp_handle(mod))); nothing)
admm_update_z(env::Env, mod::MpMod, device=nothing) =
    (mp_check(mp_handle(mod), ccall((:ea_mp_update_z, LIB), Cint, (Ptr{Cvoid}, Cdouble), mp_handle(mod), env.params.beta)); nothing)
admm_update_l(env::Env, mod::MpMod, device=nothing) =
    (mp_check(mp_handle(mod), ccall((:ea_mp_update_l, LIB), Cint, (Ptr{Cvoid}, Cdouble), mp_handle(mod), env.params.beta)); nothing)
admm_update_lz(env::Env, mod::MpMod, device=nothing) =
    (mp_check(mp_handle(mod), ccall((:ea_mp_update_lz, LIB), Cint, (Ptr{Cvoid}, Cdouble, Cdouble), mp_handle(mod),
                                    env.params.beta, env.params.MAX_MULTIPLIER)); nothing)
function admm_update_residual(env::Env, mod::MpMod, device=nothing)
    out = zeros(4)
    mp_check(mp_handle(mod), ccall((:ea_mp_update_residual, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), mp_handle(mod), out))
    mod.info.primres, mod.info.dualres, mod.info.norm_z_curr, mod.info.mismatch = out; return
end
function admm_two_level(env::Env, mod::MpMod, device=nothing)       # the whole loop in one ccall
    p = env.params
    par = EaParams(p.mu_max, p.max_auglag, p.verbose, p.initial_beta, p.inc_c, p.theta, p.outer_eps,
                   p.MAX_MULTIPLIER, p.scale, p.obj_scale, p.outer_iterlim, p.inner_iterlim)
    info = EaInfo(); err = Ref{Cdouble}(0)
    mp_check(mp_handle(mod), ccall((:ea_mp_admm_two_level, LIB), Cint, (Ptr{Cvoid}, Ref{EaParams}, Ref{EaInfo}, Ref{Cdouble}),
                                   mp_handle(mod), par, info, err))
    i = mod.info
    i.status = (:NotSpecified, :IterationLimit, :Solved)[info.status + 1]
    i.inner, i.outer, i.cumul, i.objval, i.mismatch = info.inner, info.outer, info.cumul, info.objval, info.mismatch
    i.primres, i.dualres, i.norm_z_curr, i.norm_z_prev = info.primres, info.dualres, info.norm_z_curr, info.norm_z_prev
    i.time_overall = info.time_overall; i.user.err_ramp = err[]; p.beta = info.beta
    return
end

# ---- one-level ADMM on the SQP sub-problem: ModelQpsub{Float64,TD,TI,TM} (src/models/qpsub/) --------------------
# The SQP driver fills in the host fields of the model (mod.Hs ... mod.qpsub_Qd, plain Arrays for this tag) and calls
# init_solution!, which is where they go to HBM: ea_qp_create (row-major copies: permutedims of the Julia matrices)
# + ea_qp_init_solution. The ea_qp_handle_t* is kept in the gen_solution slot.
const QpMod = ModelQpsub{Float64,TD,TI,TM}
qp_handle(mod::QpMod) = mod.gen_solution.handle
qp_check(h, rc) = rc == 0 || error(unsafe_string(ccall((:ea_qp_last_error, LIB), Cstring, (Ptr{Cvoid},), h)))
struct EaQpsubData
    Hs::Ptr{Cdouble}
    LH_1h::Ptr{Cdouble}; RH_1h::Ptr{Cdouble}; LH_1i::Ptr{Cdouble}; RH_1i::Ptr{Cdouble}
    LH_1j::Ptr{Cdouble}; RH_1j::Ptr{Cdouble}; LH_1k::Ptr{Cdouble}; RH_1k::Ptr{Cdouble}
    ls::Ptr{Cdouble}; us::Ptr{Cdouble}; line_res::Ptr{Cdouble}
    pgmax::Ptr{Cdouble}; pgmin::Ptr{Cdouble}; qgmax::Ptr{Cdouble}; qgmin::Ptr{Cdouble}; c1::Ptr{Cdouble}; c2::Ptr{Cdouble}
    Pd::Ptr{Cdouble}; Qd::Ptr{Cdouble}
end
function init_solution!(mod::QpMod, sol, rho_pq::Float64, rho_va::Float64, device=nothing)
    rm(a) = Array(permutedims(a))                       # (nline, k) column-major -> nline x k row-major
    nl = mod.grid_data.nline
    Hs = Array(permutedims(reshape(permutedims(mod.Hs), 6, 6, nl), (2, 1, 3)))    # nline blocks, each 6 x 6 row-major
    keep = (Hs, rm(mod.LH_1h), mod.RH_1h, rm(mod.LH_1i), mod.RH_1i, rm(mod.LH_1j), mod.RH_1j, rm(mod.LH_1k), mod.RH_1k,
            rm(mod.ls), rm(mod.us), Array(mod.line_res), mod.qpsub_pgmax, mod.qpsub_pgmin, mod.qpsub_qgmax,
            mod.qpsub_qgmin, mod.qpsub_c1, mod.qpsub_c2, mod.qpsub_Pd, mod.qpsub_Qd)
    GC.@preserve keep begin
        data = EaQpsubData(map(pointer, keep)...)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        g = mod.grid_data
        GC.@preserve g qp_check(C_NULL, ccall((:ea_qp_create, LIB), Cint, (Ref{EaGrid}, Ref{EaQpsubData}, Cint, Ref{Ptr{Cvoid}}),
                                              grid_struct(g), data, 0, h))
    end
    mod.gen_solution = B200Handle(h[], :ea_qp_destroy)   # finalizer -> ea_qp_destroy
    qp_check(h[], ccall((:ea_qp_init_solution, LIB), Cint, (Ptr{Cvoid}, Cdouble, Cdouble), h[], rho_pq, rho_va)); return
end
function admm_update_x(env::Env, mod::QpMod, device=nothing)
    p = env.params
    qp_check(qp_handle(mod), ccall((:ea_qp_update_x, LIB), Cint, (Ptr{Cvoid}, Int64, Int32, Cdouble, Cdouble),
                                   qp_handle(mod), mod.info.inner, p.max_auglag, p.mu_max, p.scale)); return
end
admm_update_xbar(env::Env, mod::QpMod, device=nothing) =
    (qp_check(qp_handle(mod), ccall((:ea_qp_update_xbar, LIB), Cint, (Ptr{Cvoid},), qp_handle(mod))); nothing)
admm_update_l_single(env::Env, mod::QpMod, device=nothing) =
    (qp_check(qp_handle(mod), ccall((:ea_qp_update_l_single, LIB), Cint, (Ptr{Cvoid},), qp_handle(mod))); nothing)
function admm_update_residual(env::Env, mod::QpMod, device=nothing)
    out = zeros(5)
    qp_check(qp_handle(mod), ccall((:ea_qp_update_residual, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), qp_handle(mod), out))
    mod.info.primres, mod.info.dualres, mod.info.mismatch, mod.info.objval, mod.info.auglag = out; return
end
function admm_poststep(env::Env, mod::QpMod, device=nothing)
    obj = Ref{Cdouble}(0); al = Ref{Cdouble}(0)
    qp_check(qp_handle(mod), ccall((:ea_qp_poststep, LIB), Cint,
                                   (Ptr{Cvoid}, Ref{Cdouble}, Ref{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                                   qp_handle(mod), obj, al, mod.dw_sol, mod.dtheta_sol, mod.dual_infeas))
    mod.info.objval, mod.info.auglag = obj[], al[]
    u = zeros(mod.nvar); copyto!(u, mod.solution.u_curr)
    ng, nl = mod.grid_data.ngen, mod.grid_data.nline
    mod.dpg_sol .= u[1:2:2ng]; mod.dqg_sol .= u[2:2:2ng]
    mod.dline_fl .= reshape(u[2ng+1:end], 8, nl)[1:4, :]
    qp_check(qp_handle(mod), ccall((:ea_qp_get_line_array, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Int64),
                                   qp_handle(mod), 0, mod.dline_var, 6nl)); return        # EA_QP_SQP_LINE, 6 x nline column-major
end
function admm_one_level(env::Env, mod::QpMod, device=nothing)       # the whole loop in one ccall
    p = env.params
    par = EaParams(p.mu_max, p.max_auglag, p.verbose, 0.0, p.inc_c, p.theta, p.outer_eps,
                   p.MAX_MULTIPLIER, p.scale, p.obj_scale, p.outer_iterlim, 1)
    info = EaInfo()
    qp_check(qp_handle(mod), ccall((:ea_qp_admm_one_level, LIB), Cint, (Ptr{Cvoid}, Ref{EaParams}, Ref{EaInfo}),
                                   qp_handle(mod), par, info))
    i = mod.info
    i.status = (:NotSpecified, :IterationLimit, :Solved)[info.status + 1]
    i.inner, i.outer, i.cumul, i.objval, i.auglag = info.inner, info.outer, info.cumul, info.objval, info.auglag
    i.primres, i.dualres, i.mismatch, i.time_overall = info.primres, info.dualres, info.mismatch, info.time_overall
    p.initial_beta = 0; p.beta = 0; p.inner_iterlim = 1              # admm_one_level.jl:17-22
    admm_poststep(env, mod, device)
    return
end

end # module
