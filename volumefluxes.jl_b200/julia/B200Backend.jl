# B200Backend.jl -- the reference-side binding: a new `Backend` for VolumeFluxes.jl (module SinFVM)
# whose hot path runs in libswe_b200.so (include/swe_b200.h) through `ccall`.
#
# NOT EXECUTABLE IN THE BUILD CONTAINER (no julia binary there): it documents, line by line, the stub
# a maintainer adds to the reference.  The same C ABI is exercised by the Python/ctypes host
# (volumefluxes.jl_b200/api.py), which is what the tests drive.
#
# Where it plugs in (paths relative to the reference's src/):
#   * `struct B200Backend <: Backend` next to KernelAbstractionBackend   backends/kernel_abstractions.jl:24-34
#   * `make_b200_backend`, `name`, `get_available_backends`              backends/kernel_abstractions.jl:36-55
#   * one method of `Simulator(...)`, `set_current_state!`, `perform_step!`, `simulate_to_time`,
#     `current_interior_state` for `Simulator{<:B200Backend}`            simulator.jl:72-212
#   * optional per-operator methods (`reconstruct!`, `update_bc!`, ...)   see the bottom of this file
# Include it from SinFVM.jl after simulator.jl.

const libswe = get(ENV, "SWE_B200_LIB", "libswe_b200.so")

struct B200Backend{T} <: Backend
    realtype::Type{T}          # read at system.jl:22,63; simulator.jl:80; buffer.jl:30
    strict::Bool               # bit-faithful arithmetic (Float32 = the CPU backend's mixed promotion)
    fused::Bool
    device::Int
end
make_b200_backend(::Type{T}=Float64; strict=false, fused=true, device=0) where {T} = B200Backend{T}(T, strict, fused, device)
name(::B200Backend) = "B200"
# backends/buffer.jl:23,29-31: the generic fallbacks `convert_to_backend(backend, array) = array` and
# `create_buffer(backend, nvar, size) = zeros(backend.realtype, ...)` apply to this backend unchanged, i.e. bathymetry,
# infiltration factor and `new_state` stay ordinary host arrays (the library owns all device memory and copies what
# it is handed).  `ConservedSystem(...)` (system.jl:37-53) still allocates its two face volumes and the wavespeed plane
# through them -- host memory this backend never touches; a maintainer who minds the 7 words per cell can give
# `create_volume(::B200Backend, grid, equation)` a lazily allocated variant.
# get_available_backends() (kernel_abstractions.jl:45-55): push `make_b200_backend()` when `dlopen(libswe)` succeeds.

# mirror of swe_b200_config (include/swe_b200.h); field order and types must match exactly
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

# The Simulator for this backend owns an opaque handle instead of `substep_outputs` Volumes.
mutable struct B200Simulator{T,SystemType,TimeStepperType,GridType}
    backend::B200Backend{T}
    system::SystemType
    timestepper::TimeStepperType
    grid::GridType
    handle::Ptr{Cvoid}
    cfl::T
end

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
    sim = B200Simulator{T,typeof(system),typeof(timestepper),typeof(grid)}(backend, system, timestepper, grid, h[], T(cfl))
    finalizer(s -> ccall((:swe_b200_destroy, libswe), Cint, (Ptr{Cvoid},), s.handle), sim)
    if hasB   # Julia arrays are column-major (Nx+1, Ny+1): exactly the layout the ABI expects
        Bh = Array{T}(B.B)
        check(ccall((:swe_b200_set_bathymetry, libswe), Cint, (Ptr{Cvoid}, Ptr{T}, Int64), sim.handle, Bh, length(Bh)), sim.handle)
    end
    if inf isa HortonInfiltration
        f = Array{Float64}(inf.factor)
        check(ccall((:swe_b200_set_infiltration_factor, libswe), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64), sim.handle, f, length(f)), sim.handle)
    end
    rain isa FunctionalRain && upload_functional_rain!(sim, rain, t0)     # values of the closure at t0 (see below)
    return sim
end

functional_rain(sim::B200Simulator) = (r = filter(s -> s isa FunctionalRain, sim.system.source_terms); isempty(r) ? nothing : r[1])

# simulator.jl:119-139 -- new_state is an (nx, ny) array of SVector{3} (or three scalar arrays)
function set_current_state!(sim::B200Simulator{T}, new_state) where {T}
    w = T[s[1] for s in new_state]; hu = T[s[2] for s in new_state]; hv = T[s[3] for s in new_state]
    check(ccall((:swe_b200_set_state_interior, libswe), Cint, (Ptr{Cvoid}, Ptr{T}, Ptr{T}, Ptr{T}), sim.handle, w, hu, hv), sim.handle)
end

# simulator.jl:114-115 + interior_volume_variable.jl:168-176: collect(current_interior_state(sim).h) etc.
function current_interior_state(sim::B200Simulator{T}) where {T}
    n = interior_size(sim.grid)
    w = Array{T}(undef, n); hu = similar(w); hv = similar(w)
    check(ccall((:swe_b200_get_state_interior, libswe), Cint, (Ptr{Cvoid}, Ptr{T}, Ptr{T}, Ptr{T}), sim.handle, w, hu, hv), sim.handle)
    return (h=w, hu=hu, hv=hv)
end

function time_and_step(sim::B200Simulator)
    t = Ref(0.0); dt = Ref(0.0)
    check(ccall((:swe_b200_get_time, libswe), Cint, (Ptr{Cvoid}, Ref{Float64}, Ref{Float64}), sim.handle, t, dt), sim.handle)
    return t[], dt[]
end
current_time(sim::B200Simulator) = time_and_step(sim)[1]
current_timestep(sim::B200Simulator) = time_and_step(sim)[2]

# simulator.jl:145-177
function perform_step!(sim::B200Simulator, max_dt)
    fr = functional_rain(sim)       # a (possibly time-dependent) closure: re-evaluate it at the time of this step
    isnothing(fr) || upload_functional_rain!(sim, fr, current_time(sim))
    check(ccall((:swe_b200_step, libswe), Cint, (Ptr{Cvoid}, Float64, Cint), sim.handle, max_dt, 0), sim.handle)
end

# simulator.jl:179-212.  Without callback: device-resident loop; with callback: the reference's host loop.
function simulate_to_time(sim::B200Simulator, endtime; match_endtime=true, callback=nothing,
                          show_progress=true, maximum_timestep=nothing)
    fr = functional_rain(sim)
    isnothing(fr) || isnothing(callback) && (callback = (t, s) -> nothing)   # the closure needs the host between steps
    if isnothing(callback)
        n = Ref{Int64}(0)
        check(ccall((:swe_b200_simulate_to_time, libswe), Cint, (Ptr{Cvoid}, Float64, Cint, Float64, Ref{Int64}),
                    sim.handle, endtime, match_endtime ? 1 : 0, isnothing(maximum_timestep) ? 0.0 : maximum_timestep, n), sim.handle)
        return n[]
    end
    Base.isiterable(typeof(callback)) && (callback = MultipleCallbacks(callback))
    t = current_time(sim)
    while t < endtime
        max_dt = match_endtime ? endtime - t : Inf
        isnothing(maximum_timestep) || (max_dt = min(max_dt, maximum_timestep))
        isnothing(fr) || upload_functional_rain!(sim, fr, t)
        check(ccall((:swe_b200_step, libswe), Cint, (Ptr{Cvoid}, Float64, Cint), sim.handle, max_dt, 1), sim.handle)
        t = current_time(sim)
        callback(t, sim)
    end
end

# ---- optional per-operator dispatch points (bc.jl:113-114, linearreconstruction.jl:58, system.jl:62) ------------
update_bc!(sim::B200Simulator, buffer::Integer=0) =
    check(ccall((:swe_b200_update_bc, libswe), Cint, (Ptr{Cvoid}, Cint), sim.handle, buffer), sim.handle)
reconstruct!(sim::B200Simulator, direction::Direction, buffer::Integer=0) =
    check(ccall((:swe_b200_reconstruct, libswe), Cint, (Ptr{Cvoid}, Cint, Cint), sim.handle, Base.to_index(direction) - 1, buffer), sim.handle)
function add_time_derivative!(sim::B200Simulator, in_buffer::Integer, out_buffer::Integer, t)
    ms = zeros(Float64, 2)
    check(ccall((:swe_b200_add_time_derivative, libswe), Cint, (Ptr{Cvoid}, Cint, Cint, Float64, Ptr{Float64}),
                sim.handle, in_buffer, out_buffer, t, ms), sim.handle)
    return ms
end

# ---- FunctionalRain (source_term_rain.jl:57-68): the closure is evaluated on the host for all cells (cell_center,
# grid.jl:117-121) and uploaded; call before every step for a time-dependent rain_function ----------------------
function upload_functional_rain!(sim::B200Simulator, rain::FunctionalRain, t)
    g = sim.grid
    vals = [Float64(rain.rain_function(t, cell_center(g, CartesianIndex(i, j))...)) for i in 1:g.totalcells[1], j in 1:g.totalcells[2]]
    check(ccall((:swe_b200_set_rain_plane, libswe), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64), sim.handle, vals, length(vals)), sim.handle)
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
begin_snapshot!(sim::B200Simulator{T}, w::Array{T}, hu::Array{T}, hv::Array{T}) where {T} =
    check(ccall((:swe_b200_snapshot_begin, libswe), Cint, (Ptr{Cvoid}, Ptr{T}, Ptr{T}, Ptr{T}), sim.handle, w, hu, hv), sim.handle)
wait_snapshot!(sim::B200Simulator) = check(ccall((:swe_b200_snapshot_wait, libswe), Cint, (Ptr{Cvoid},), sim.handle), sim.handle)

# ---- scalar reductions for callbacks (examples/example_tools.jl:62-90 TotalWaterVolume collects the whole state;
# here sum(w - B_cell) etc. are reduced on the device and 64 bytes come back) ------------------------------------
function diagnostics(sim::B200Simulator, buffer::Integer=0)
    out = zeros(Float64, 8)
    check(ccall((:swe_b200_diagnostics, libswe), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), sim.handle, buffer, out), sim.handle)
    return (sum_h=out[1], sum_hu=out[2], sum_hv=out[3], max_h=out[4], min_h=out[5], wet_cells=Int(out[6]),
            max_momentum=out[7], sum_w=out[8])
end
total_water_volume(sim::B200Simulator) = diagnostics(sim).sum_h

# ---- the kernel surface: closures cannot run inside a precompiled kernel, so the iteration space is
# enumerated on the device and `f` is applied on the host (kernel_abstractions.jl:66-131) ----------------------
function for_each_inner_cell(f, sim::B200Simulator, direction; ghostcells=sim.grid.ghostcells[direction])
    n = Ref{Int64}(0); d = Base.to_index(direction) - 1
    check(ccall((:swe_b200_enumerate_inner_cells, libswe), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Int64}, Ref{Int64}), sim.handle, d, ghostcells, C_NULL, n))
    idx = Matrix{Int64}(undef, 6, n[])
    check(ccall((:swe_b200_enumerate_inner_cells, libswe), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Int64}, Ref{Int64}), sim.handle, d, ghostcells, idx, n))
    for c in eachcol(idx)
        f(CartesianIndex(c[1], c[2]), CartesianIndex(c[3], c[4]), CartesianIndex(c[5], c[6]))
    end
end
