# B200Backend.jl -- the reference-side binding: a new `Backend` for VolumeFluxes.jl (module VolumeFluxes, formerly
# SinFVM) whose hot path runs in libswe_b200.so (include/swe_b200.h) through `ccall`.
#
# NOT EXECUTABLE IN THE BUILD CONTAINER (no julia binary there): it is the stub a maintainer adds to the reference.
# The same C ABI is exercised by the Python/ctypes host (volumefluxes.jl_b200/api.py), which is what the tests drive;
# tests/test_julia_shim.py checks this file against the header (struct layout, ccall symbols and arities).
#
# Where it plugs in (paths relative to the reference's src/):
#   * `struct B200Backend <: Backend` next to KernelAbstractionBackend        backends/kernel_abstractions.jl:24-34
#   * `make_b200_backend`, `name`, `get_available_backends`                   backends/kernel_abstractions.jl:36-55
#   * `create_buffer` / `create_scalar` / `convert_to_backend` methods        backends/buffer.jl:23-42, system.jl:21-22
#   * `for_each_cell` / `for_each_inner_cell` / `for_each_ghost_cell` methods backends/kernel_abstractions.jl:73, 83, 129
#   * `Simulator(backend::B200Backend, ...)` builds a genuine `Simulator{<:B200Backend,...}` (simulator.jl:60-91), so
#     every reference method typed `::Simulator` (current_time, current_timestep, callbacks.jl:49-55 ...) applies;
#     `set_current_state!`, `current_interior_state`, `perform_step!`, `simulate_to_time` get methods for
#     `Simulator{<:B200Backend}`                                              simulator.jl:98-212
#   * optional per-operator methods (`reconstruct!`, `update_bc!`, ...)       see the bottom of this file
# Include it from VolumeFluxes.jl after simulator.jl.

const libswe = get(ENV, "SWE_B200_LIB", "libswe_b200.so")

struct B200Backend{T} <: Backend
    realtype::Type{T}          # read at system.jl:22,63; simulator.jl:80; buffer.jl:30
    strict::Bool               # bit-faithful arithmetic (Float32 = the CPU backend's mixed promotion)
    fused::Bool
    device::Int
end
make_b200_backend(::Type{T}=Float64; strict=false, fused=true, device=0) where {T} = B200Backend{T}(T, strict, fused, device)
name(::B200Backend) = "B200"
# get_available_backends() (kernel_abstractions.jl:45-55): push `make_b200_backend()` when `dlopen(libswe)` succeeds.

# ---- buffers (backends/buffer.jl:23-42, system.jl:21-22) --------------------------------------------------------
# The library owns all device memory: the three RK volumes live inside the handle, the face volumes / wavespeed plane
# only exist when a per-operator entry point is used.  `ConservedSystem(...)` (system.jl:37-53) nevertheless calls
# create_volume twice and create_scalar once; with the generic fallbacks that would be 7 host words per cell nobody
# reads (7.5 GB at 16384^2 Float32).  These methods hand it zero-cost placeholders instead.
struct LazyZeros{T,N} <: AbstractArray{T,N}      # reads as zeros, owns no memory, cannot be written
    dims::NTuple{N,Int}
end
Base.size(a::LazyZeros) = a.dims
Base.getindex(a::LazyZeros{T}, I::Vararg{Int}) where {T} = zero(T)
Base.setindex!(a::LazyZeros, v, I::Vararg{Int}) =
    error("B200Backend buffers are device-resident: use set_current_state! / the swe_b200_* operators")
create_buffer(backend::B200Backend{T}, number_of_variables::Int64, spatial_resolution) where {T} =
    LazyZeros{T,length(spatial_resolution) + 1}((spatial_resolution..., number_of_variables))
create_scalar(backend::B200Backend{T}, grid, equation) where {T} = LazyZeros{T,length(size(grid))}(Tuple(size(grid)))
convert_to_backend(backend::B200Backend, array::AbstractArray) = array     # host arrays are copied by the library

# ---- mirror of swe_b200_config (include/swe_b200.h); field order and types must match exactly -------------------
struct SweB200Config
    abi_version::Int32; dtype::Int32; nx::Int32; ny::Int32; gc::Int32; bc::Int32; eq_kind::Int32
    recon::Int32; stepper::Int32; friction::Int32; has_bathymetry::Int32; has_bottom_source::Int32
    rain_kind::Int32; infil_kind::Int32; infil_first::Int32; n_rain::Int32; strict::Int32; fused::Int32
    device::Int32; rank::Int32; nranks::Int32; global_ny::Int32; y_offset::Int32
    reserved::NTuple{3,Int32}
    extent::NTuple{4,Float64}
    theta::Float64; rho::Float64; g::Float64; depth_cutoff::Float64; kappa::Float64; cfl::Float64
    Cz::Float64; const_B::Float64; t0::Float64
    rain_rate::Float64; infil_const::Float64; f0::Float64; fc::Float64; k::Float64
    rain_rates::NTuple{16,Float64}; rain_times::NTuple{16,Float64}
end

bc_code(::PeriodicBC) = Int32(0); bc_code(::WallBC) = Int32(1); bc_code(::NeumannBC) = Int32(2)
friction_code(::Nothing) = Int32(0)
friction_code(f::ImplicitFriction) = f.friction_function === friction_bh2021 ? Int32(1) :
                                     f.friction_function === friction_bsa2012 ? Int32(2) : Int32(3)

function check(rc, h=C_NULL)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:swe_b200_last_error, libswe), Cstring, (Ptr{Cvoid},), h))
    rc == -2 && throw(DomainError(msg))          # bottom_topography.jl:78-82, source_term_infiltration.jl:61-65
    rc == -1 && throw(ArgumentError(msg))        # equation.jl:83, linearreconstruction.jl:59
    error("swe_b200 error $rc: $msg")
end

# ---- the Simulator of this backend IS a `Simulator{...}` (simulator.jl:60-70).  Its substep_outputs are not Volumes
# but views of the library's buffers 0..nsub: (handle, buffer index).  The first one owns the handle. -------------
mutable struct B200State{T}
    handle::Ptr{Cvoid}
    buffer::Int32
end
handle(sim::Simulator{<:B200Backend}) = sim.substep_outputs[1].handle

function Simulator(backend::B200Backend{T}, system, timestepper, grid::CartesianGrid{2}; cfl=0.25, t0=0.0) where {T}
    eq = system.equation
    B = eq isa AllPracticalSWE ? eq.B : ConstantBottomTopography()
    hasB = B isa BottomTopography2D && !is_zero(B)
    rains = filter(s -> s isa SourceTermRain, system.source_terms)
    infs = filter(s -> s isa SourceTermInfiltration, system.source_terms)
    rain = isempty(rains) ? nothing : rains[1]
    inf = isempty(infs) ? nothing : infs[1]
    infil_first = (rain !== nothing && inf !== nothing &&
                   findfirst(==(inf), system.source_terms) < findfirst(==(rain), system.source_terms))
    rates = ntuple(i -> rain isa TimeDependentRain && i <= length(rain.rain_rates) ? Float64(rain.rain_rates[i]) : 0.0, 16)
    times = ntuple(i -> rain isa TimeDependentRain && i <= length(rain.time) ? Float64(rain.time[i]) : 0.0, 16)
    fr = system.implicit_source_term
    cfg = Ref(SweB200Config(1, T === Float32 ? 0 : 1, interior_size(grid)..., grid.ghostcells[1], bc_code(grid.boundary),
        eq isa AllPureSWE ? 1 : 0, system.reconstruction isa LinearReconstruction ? 1 : 0,
        timestepper isa RungeKutta2 ? 1 : 0, friction_code(fr), hasB ? 1 : 0,
        any(s -> s isa SourceTermBottom, system.source_terms) ? 1 : 0,
        rain isa ConstantRain ? 1 : rain isa TimeDependentRain ? 2 : rain isa FunctionalRain ? 3 : 0,
        inf isa ConstantInfiltration ? 1 : inf isa HortonInfiltration ? 2 : 0, infil_first ? 1 : 0,
        rain isa TimeDependentRain ? length(rain.rain_rates) : 0, backend.strict ? 1 : 0, backend.fused ? 1 : 0,
        backend.device, 0, 1, interior_size(grid)[2], 0, (Int32(0), Int32(0), Int32(0)),
        Tuple(grid.extent),
        system.reconstruction isa LinearReconstruction ? system.reconstruction.theta : 1.0,
        Float64(eq.ρ), Float64(eq.g),
        eq isa AllPracticalSWE ? Float64(eq.depth_cutoff) : 1e-5, eq isa AllPracticalSWE ? Float64(eq.desingularizing_kappa) : 1e-5,
        Float64(cfl), fr === nothing ? 0.0 : Float64(fr.Cz), hasB ? 0.0 : Float64(B isa ConstantBottomTopography ? B.B : 0.0),
        Float64(t0), rain isa ConstantRain ? Float64(rain.rain_rate) : 0.0,
        inf isa ConstantInfiltration ? Float64(inf.infiltration_rate) : 0.0,
        inf isa HortonInfiltration ? Float64(inf.f0) : 0.0, inf isa HortonInfiltration ? Float64(inf.fc) : 0.0,
        inf isa HortonInfiltration ? Float64(inf.k) : 0.0, rates, times))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:swe_b200_create, libswe), Cint, (Ref{SweB200Config}, Ref{Ptr{Cvoid}}), cfg, h))
    states = [B200State{T}(h[], Int32(k)) for k in 0:number_of_substeps(timestepper)]
    finalizer(s -> ccall((:swe_b200_destroy, libswe), Cint, (Ptr{Cvoid},), s.handle), states[1])
    # the default constructor of the reference's struct (simulator.jl:60-70): same fields, same order
    sim = Simulator{typeof(backend),typeof(system),typeof(timestepper),typeof(grid),B200State{T},T}(
        backend, system, timestepper, grid, states, MVector{1,T}([0]), T(cfl), MVector{1,T}([t0]))
    if hasB   # Julia arrays are column-major (Nx+1, Ny+1): exactly the layout the ABI expects
        Bh = Array{T}(B.B)
        check(ccall((:swe_b200_set_bathymetry, libswe), Cint, (Ptr{Cvoid}, Ptr{T}, Int64), handle(sim), Bh, length(Bh)), handle(sim))
    end
    if inf isa HortonInfiltration
        f = Array{Float64}(inf.factor)
        check(ccall((:swe_b200_set_infiltration_factor, libswe), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64), handle(sim), f, length(f)), handle(sim))
    end
    rain isa FunctionalRain && upload_functional_rain!(sim, rain, t0)     # values of the closure at t0 (see below)
    return sim
end

functional_rain(sim::Simulator{<:B200Backend}) =
    (r = filter(s -> s isa FunctionalRain, sim.system.source_terms); isempty(r) ? nothing : r[1])

# which kernel family serves this simulator: 0 per-operator, 1 fused generic, 2 fused x2 (Float32), 3 fused d1 (Float64)
kernel_family(sim::Simulator{<:B200Backend}) = ccall((:swe_b200_kernel_family, libswe), Cint, (Ptr{Cvoid},), handle(sim))

# simulator.jl:119-139 -- new_state is an (nx, ny) array of SVector{3} (or three scalar arrays)
function set_current_state!(sim::Simulator{B200Backend{T}}, new_state) where {T}
    w = T[s[1] for s in new_state]; hu = T[s[2] for s in new_state]; hv = T[s[3] for s in new_state]
    check(ccall((:swe_b200_set_state_interior, libswe), Cint, (Ptr{Cvoid}, Ptr{T}, Ptr{T}, Ptr{T}), handle(sim), w, hu, hv), handle(sim))
end

# simulator.jl:114-115 + interior_volume_variable.jl:168-176: collect(current_interior_state(sim).h) etc.
function current_interior_state(sim::Simulator{B200Backend{T}}) where {T}
    n = interior_size(sim.grid)
    w = Array{T}(undef, n); hu = similar(w); hv = similar(w)
    check(ccall((:swe_b200_get_state_interior, libswe), Cint, (Ptr{Cvoid}, Ptr{T}, Ptr{T}, Ptr{T}), handle(sim), w, hu, hv), handle(sim))
    return (h=w, hu=hu, hv=hv)
end

# dt and t live on the device; the reference keeps them in the simulator's MVectors (simulator.jl:66,68,141-142), which
# `current_time` / `current_timestep` / the callbacks read: refresh the mirrors whenever the host gets control back
function refresh_time!(sim::Simulator{<:B200Backend})
    t = Ref(0.0); dt = Ref(0.0)
    check(ccall((:swe_b200_get_time, libswe), Cint, (Ptr{Cvoid}, Ref{Float64}, Ref{Float64}), handle(sim), t, dt), handle(sim))
    sim.t[1] = t[]; sim.current_timestep[1] = dt[]
    return t[], dt[]
end

# simulator.jl:145-177 (does not advance t, like the reference: simulate_to_time does `t[1] += dt`)
function perform_step!(sim::Simulator{<:B200Backend}, max_dt)
    fr = functional_rain(sim)       # a (possibly time-dependent) closure: re-evaluate it at the time of this step
    isnothing(fr) || upload_functional_rain!(sim, fr, sim.t[1])
    check(ccall((:swe_b200_step, libswe), Cint, (Ptr{Cvoid}, Float64, Cint), handle(sim), max_dt, 0), handle(sim))
    refresh_time!(sim)
end

# simulator.jl:179-212.  Without callback: device-resident loop; with callback: the reference's host loop.
function simulate_to_time(sim::Simulator{<:B200Backend}, endtime; match_endtime=true, callback=nothing,
                          show_progress=true, maximum_timestep=nothing)
    fr = functional_rain(sim)
    isnothing(fr) || isnothing(callback) && (callback = (t, s) -> nothing)   # the closure needs the host between steps
    if isnothing(callback)
        n = Ref{Int64}(0)
        check(ccall((:swe_b200_simulate_to_time, libswe), Cint, (Ptr{Cvoid}, Float64, Cint, Float64, Ref{Int64}),
                    handle(sim), endtime, match_endtime ? 1 : 0, isnothing(maximum_timestep) ? 0.0 : maximum_timestep, n), handle(sim))
        refresh_time!(sim)
        return n[]
    end
    Base.isiterable(typeof(callback)) && (callback = MultipleCallbacks(callback))
    t = sim.t[1]
    while t < endtime
        max_dt = match_endtime ? endtime - t : Inf
        isnothing(maximum_timestep) || (max_dt = min(max_dt, maximum_timestep))
        isnothing(fr) || upload_functional_rain!(sim, fr, t)
        check(ccall((:swe_b200_step, libswe), Cint, (Ptr{Cvoid}, Float64, Cint), handle(sim), max_dt, 1), handle(sim))
        t, _ = refresh_time!(sim)
        callback(t, sim)
    end
end

# ---- optional per-operator dispatch points (bc.jl:113-114, linearreconstruction.jl:58, system.jl:62) ------------
update_bc!(sim::Simulator{<:B200Backend}, state::B200State) =
    check(ccall((:swe_b200_update_bc, libswe), Cint, (Ptr{Cvoid}, Cint), state.handle, state.buffer), state.handle)
reconstruct!(sim::Simulator{<:B200Backend}, direction::Direction, state::B200State=sim.substep_outputs[1]) =
    check(ccall((:swe_b200_reconstruct, libswe), Cint, (Ptr{Cvoid}, Cint, Cint), state.handle, Base.to_index(direction) - 1, state.buffer), state.handle)
function add_time_derivative!(output::B200State, sim::Simulator{<:B200Backend}, current_state::B200State, t)
    ms = zeros(Float64, 2)
    check(ccall((:swe_b200_add_time_derivative, libswe), Cint, (Ptr{Cvoid}, Cint, Cint, Float64, Ptr{Float64}),
                output.handle, current_state.buffer, output.buffer, t, ms), output.handle)
    return ms
end
implicit_substep!(output::B200State, previous::B200State, sim::Simulator{<:B200Backend}, dt) =
    check(ccall((:swe_b200_implicit_substep, libswe), Cint, (Ptr{Cvoid}, Cint, Cint, Float64), output.handle, output.buffer, previous.buffer, dt), output.handle)
post_proc_substep!(output::B200State, sim::Simulator{<:B200Backend}) =
    check(ccall((:swe_b200_post_proc_substep, libswe), Cint, (Ptr{Cvoid}, Cint), output.handle, output.buffer), output.handle)

# ---- FunctionalRain (source_term_rain.jl:57-68): the closure is evaluated on the host for all cells (cell_center,
# grid.jl:117-121) and uploaded; call before every step for a time-dependent rain_function ----------------------
function upload_functional_rain!(sim::Simulator{<:B200Backend}, rain::FunctionalRain, t)
    g = sim.grid
    vals = [Float64(rain.rain_function(t, cell_center(g, CartesianIndex(i, j))...)) for i in 1:g.totalcells[1], j in 1:g.totalcells[2]]
    check(ccall((:swe_b200_set_rain_plane, libswe), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64), handle(sim), vals, length(vals)), handle(sim))
end

# ---- asynchronous snapshots for writer callbacks (callbacks.jl:40-55, examples/example_tools.jl:113-137): the copy to
# the (page-locked) host arrays overlaps the following steps; wait_snapshot! before reading them ------------------
function pinned_array(::Type{T}, dims...) where {T}
    p = ccall((:swe_b200_host_alloc, libswe), Ptr{Cvoid}, (Int64,), sizeof(T) * prod(dims))
    p == C_NULL && error("cudaHostAlloc failed")
    a = unsafe_wrap(Array, Ptr{T}(p), dims; own=false)
    finalizer(_ -> ccall((:swe_b200_host_free, libswe), Cint, (Ptr{Cvoid},), p), a)
    return a
end
begin_snapshot!(sim::Simulator{B200Backend{T}}, w::Array{T}, hu::Array{T}, hv::Array{T}) where {T} =
    check(ccall((:swe_b200_snapshot_begin, libswe), Cint, (Ptr{Cvoid}, Ptr{T}, Ptr{T}, Ptr{T}), handle(sim), w, hu, hv), handle(sim))
wait_snapshot!(sim::Simulator{<:B200Backend}) = check(ccall((:swe_b200_snapshot_wait, libswe), Cint, (Ptr{Cvoid},), handle(sim)), handle(sim))

# ---- scalar reductions for callbacks (examples/example_tools.jl:62-90 TotalWaterVolume collects the whole state;
# here sum(w - B_cell) etc. are reduced on the device and 64 bytes come back) ------------------------------------
function diagnostics(sim::Simulator{<:B200Backend}, buffer::Integer=0)
    out = zeros(Float64, 8)
    check(ccall((:swe_b200_diagnostics, libswe), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), handle(sim), buffer, out), handle(sim))
    return (sum_h=out[1], sum_hu=out[2], sum_hv=out[3], max_h=out[4], min_h=out[5], wet_cells=Int(out[6]),
            max_momentum=out[7], sum_w=out[8])
end
total_water_volume(sim::Simulator{<:B200Backend}) = diagnostics(sim).sum_h

# ---- the kernel surface (backends/kernel_abstractions.jl:73, 83, 129), same signatures as the reference ----------
# A precompiled backend cannot run an arbitrary Julia closure inside a kernel (north_star: no KernelAbstractions /
# CUDA.jl codegen).  The hot-path CALLERS of this surface (reconstruct!, compute_flux!, do_substep!, update_bc!, ...)
# never reach it on this backend: perform_step! above replaces them by the fused kernels.  What remains are utility
# callers on host arrays (volume.jl:147-163 setindex loops etc.): for those the index space is enumerated by the
# library's device kernel (swe_b200_enumerate_grid, bit-for-bit the spaces of grid.jl:213-253) and `f` runs on the
# host over it, with the trailing arguments y... handed through like the reference does.
function enumerate_grid(kind::Integer, backend::B200Backend, grid::CartesianGrid{2}, direction::Integer, ghostcells::Integer)
    nx, ny = interior_size(grid)
    n = Ref{Int64}(0)
    args = (Cint(kind), Cint(nx), Cint(ny), Cint(grid.ghostcells[1]), Cint(direction - 1), Cint(ghostcells), Cint(backend.device))
    check(ccall((:swe_b200_enumerate_grid, libswe), Cint, (Cint, Cint, Cint, Cint, Cint, Cint, Cint, Ptr{Int64}, Ref{Int64}), args..., C_NULL, n))
    idx = Matrix{Int64}(undef, kind == 1 ? 6 : 2, n[])
    check(ccall((:swe_b200_enumerate_grid, libswe), Cint, (Cint, Cint, Cint, Cint, Cint, Cint, Cint, Ptr{Int64}, Ref{Int64}), args..., idx, n))
    return idx
end
function for_each_cell(f, backend::B200Backend, grid, y...)
    for c in eachcol(enumerate_grid(0, backend, grid, 1, 0))
        f(CartesianIndex(c[1], c[2]), y...)
    end
end
function for_each_inner_cell(f, backend::B200Backend, grid, direction, y...; ghostcells=grid.ghostcells[direction])
    for c in eachcol(enumerate_grid(1, backend, grid, Base.to_index(direction), ghostcells))
        f(CartesianIndex(c[1], c[2]), CartesianIndex(c[3], c[4]), CartesianIndex(c[5], c[6]), y...)
    end
end
function for_each_ghost_cell(f, backend::B200Backend, grid, direction, y...)
    for c in eachcol(enumerate_grid(2, backend, grid, Base.to_index(direction), 0))
        f(CartesianIndex(c[1], c[2]), y...)
    end
end
