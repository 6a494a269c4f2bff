# dump_reference.jl -- produces REFERENCE-MADE golden vectors with the unmodified VolumeFluxes.jl package.
#
# The build container has no Julia, so tests/golden/trajectories.npz holds outputs of the CPU restatement (oracle/);
# parity is therefore "unpinned" in the strict sense (DESIGN.md section 2).  This script closes that gap on the first
# machine that has Julia + the reference checked out:
#
#   python tests/golden/make_golden.py --export-inputs /tmp/golden_io      # parameters + raw input arrays
#   julia --project=/path/to/VolumeFluxes.jl tests/golden/dump_reference.jl /tmp/golden_io
#   python tests/golden/make_golden.py --import-reference /tmp/golden_io   # -> tests/golden/reference_trajectories.npz
#
# tests/test_golden.py then checks the oracle AND the CUDA path against reference_trajectories.npz (bit-exact for
# the oracle / strict policies, north_star tolerances for the fast policies) whenever that file exists.
#
# Per case <key> (e.g. lake_bh2021.float32) the input directory holds <key>.json (parameters, see make_golden.py)
# and raw little-endian arrays <key>.w.bin / .hu.bin / .hv.bin (nx x ny, x fastest = Julia column-major) and, when the
# case has a bathymetry array, <key>.B.bin ((nx+5) x (ny+5)).  Output: <key>.ref_state.bin (nx x ny x 3, interior,
# realtype) and <key>.ref_dt.bin (NSTEPS Float64).
#
# Every case is run the way the reference's own CPU tests run: make_cpu_backend(T) (backends/kernel_abstractions.jl:38)
# with arrays of type T and the default Float64 scalars (SURVEY App. A.9), NSTEPS x { perform_step!; t += dt }
# (simulator.jl:145-177, 199-205).
using VolumeFluxes, StaticArrays, JSON

function read_array(path, T, dims)
    a = Array{T}(undef, dims...)
    read!(path, a)
    return a
end

function run_case(dir, key)
    p = JSON.parsefile(joinpath(dir, key * ".json"))
    T = p["dtype"] == "float32" ? Float32 : Float64
    nx, ny, gc = p["nx"], p["ny"], p["gc"]
    backend = VolumeFluxes.make_cpu_backend(T)
    bc = p["bc"] == "wall" ? VolumeFluxes.WallBC() : p["bc"] == "periodic" ? VolumeFluxes.PeriodicBC() : VolumeFluxes.NeumannBC()
    e = p["extent"]
    grid = VolumeFluxes.CartesianGrid(nx, ny; gc=gc, boundary=bc, extent=[e[1] e[2]; e[3] e[4]])
    sources = VolumeFluxes.SourceTerm[]
    if p["has_B"]
        B = read_array(joinpath(dir, key * ".B.bin"), T, (nx + 2gc + 1, ny + 2gc + 1))
        bottom = VolumeFluxes.BottomTopography2D(B, backend, grid)
        eq = VolumeFluxes.ShallowWaterEquations(bottom; depth_cutoff=p["depth_cutoff"], desingularizing_kappa=p["kappa"])
    else
        eq = VolumeFluxes.ShallowWaterEquations(; depth_cutoff=p["depth_cutoff"], desingularizing_kappa=p["kappa"])
    end
    # source-term list order as in the examples (terrain.jl:107-111): infiltration, rain, bottom
    inf = p["infiltration"]
    if inf == "horton"
        push!(sources, VolumeFluxes.HortonInfiltration(grid, backend))
    elseif inf isa Number
        push!(sources, VolumeFluxes.ConstantInfiltration(inf))
    end
    isnothing(p["rain"]) || push!(sources, VolumeFluxes.ConstantRain(p["rain"]))
    p["bottom_source"] && push!(sources, VolumeFluxes.SourceTermBottom())
    rec = p["recon"] == "linear" ? VolumeFluxes.LinearReconstruction(p["theta"]) : VolumeFluxes.NoReconstruction()
    laws = Dict("bh2021" => VolumeFluxes.friction_bh2021, "bsa2012" => VolumeFluxes.friction_bsa2012, "fcg2016" => VolumeFluxes.friction_fcg2016)
    friction = isnothing(p["friction"]) ? nothing : VolumeFluxes.ImplicitFriction(; friction_function=laws[p["friction"]])
    system = VolumeFluxes.ConservedSystem(backend, rec, VolumeFluxes.CentralUpwind(eq), eq, grid, sources, friction)
    stepper = p["stepper"] == "rk2" ? VolumeFluxes.RungeKutta2() : VolumeFluxes.ForwardEulerStepper()
    sim = VolumeFluxes.Simulator(backend, system, stepper, grid; cfl=p["cfl"])
    w = read_array(joinpath(dir, key * ".w.bin"), T, (nx, ny))
    hu = read_array(joinpath(dir, key * ".hu.bin"), T, (nx, ny))
    hv = read_array(joinpath(dir, key * ".hv.bin"), T, (nx, ny))
    VolumeFluxes.set_current_state!(sim, [SVector{3,T}(w[i, j], hu[i, j], hv[i, j]) for i in 1:nx, j in 1:ny])
    max_dt = isnothing(p["maximum_timestep"]) ? Inf : p["maximum_timestep"]
    dts = Float64[]
    for _ in 1:p["nsteps"]
        VolumeFluxes.perform_step!(sim, max_dt)
        sim.t[1] += sim.current_timestep[1]              # the loop body of simulate_to_time (simulator.jl:204-205)
        push!(dts, Float64(sim.current_timestep[1]))
    end
    st = VolumeFluxes.current_interior_state(sim)
    out = Array{T}(undef, nx, ny, 3)
    out[:, :, 1] = collect(st.h); out[:, :, 2] = collect(st.hu); out[:, :, 3] = collect(st.hv)
    write(joinpath(dir, key * ".ref_state.bin"), out)
    write(joinpath(dir, key * ".ref_dt.bin"), dts)
    println(key, ": t = ", sim.t[1], "  dt[end] = ", dts[end])
end

dir = ARGS[1]
for f in sort(readdir(dir))
    endswith(f, ".json") && run_case(dir, f[1:end-5])
end
