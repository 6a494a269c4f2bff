"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU, exports every
symbol include/swe_b200.h declares, and the ctypes mirror of swe_b200_config has the C layout."""
import ctypes as C
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import cases

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "swe_b200.h"


def test_library_loads_and_exports_all_declared_symbols(pkg):
    lib = pkg.api._lib.lib()
    declared = set(re.findall(r"\b(swe_b200_[a-z0-9_]+)\s*\(", HEADER.read_text()))
    declared -= {"swe_b200_handle", "swe_b200_config", "swe_b200_timing"}
    assert len(declared) >= 29
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/swe_b200.h but not exported"
    assert declared == set(pkg.api._lib.SYMBOLS), "ctypes binding table and header disagree"


def test_config_layout_matches_header(pkg, tmp_path):
    src = tmp_path / "layout.c"
    fields = ["abi_version", "gc", "fused", "device", "y_offset", "extent", "theta", "t0", "rain_rate", "k",
              "rain_rates", "rain_times"]
    body = "".join(f'printf("{f} %zu\\n", offsetof(swe_b200_config, {f}));\n' for f in fields)
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "swe_b200.h"\nint main(void){\n'
                   'printf("sizeof %zu\\n", sizeof(swe_b200_config));\n'
                   'printf("timing %zu\\n", sizeof(swe_b200_timing));\n' + body + "return 0;}\n")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    Cfg = pkg.api._lib.Config
    assert int(out["sizeof"]) == C.sizeof(Cfg)
    assert int(out["timing"]) == C.sizeof(pkg.api._lib.Timing)
    for f in fields:
        assert int(out[f]) == getattr(Cfg, f).offset, f


def test_header_is_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "swe_b200.h"\nint main(void){return SWE_B200_ABI_VERSION == 1 ? 0 : 1;}\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", str(ROOT / "include"), str(src), "-o",
                    str(tmp_path / "t")], check=True)


def test_host_side_validation_needs_no_gpu(pkg):
    P = pkg
    be = P.make_b200_backend(np.float32)
    assert P.name(be) == "B200" and be.realtype is np.float32
    grid = P.CartesianGrid(64, 32, gc=2, boundary=P.WallBC(), extent=[[0, 10.0], [0, 5.0]])
    assert grid.totalcells == (68, 36) and P.compute_dx(grid) == 10.0 / 64 and P.compute_dy(grid) == 5.0 / 32
    assert P.inner_cells(grid, P.XDIR) == (64, 32) and P.inner_cells(grid, P.XDIR, 0) == (68, 32)
    assert P.inner_cells(grid, P.YDIR, 1) == (64, 34)
    assert P.cell_faces(grid, interior=False).shape == (37, 69, 2)
    with pytest.raises(P.DomainError):
        P.BottomTopography2D(np.zeros((36, 68)), be, grid)
    bt = P.BottomTopography2D(np.ones((37, 69)), be, grid)
    eq = P.ShallowWaterEquations(bt, depth_cutoff=1e-3)
    with pytest.raises(P.ArgumentError):
        P.ConservedSystem(be, P.LinearReconstruction(), P.CentralUpwind(eq), eq, grid)
    P.ConservedSystem(be, P.LinearReconstruction(), P.CentralUpwind(eq), eq, grid, [P.SourceTermBottom()])
    with pytest.raises(P.DomainError):
        P.TimeDependentRain([1.0, 2.0], [0.0])
    # FunctionalRain: the closure is evaluated on the host (vectorised, or cell by cell when it branches on x / y)
    fr = P.FunctionalRain(lambda t, x, y: 2.0 * x + y if t > 0 else 0.0, grid)
    v = fr.values(1.0, np.arange(3.0), np.arange(2.0))
    assert v.shape == (2, 3) and v[1, 2] == 5.0 and np.all(fr.values(0.0, np.arange(3.0), np.arange(2.0)) == 0.0)
    assert np.array_equal(P.FunctionalRain(lambda t, x, y: 1.0 if x > 0.5 else 0.0, grid).values(0.0, np.arange(3.0), np.arange(2.0)),
                          np.array([[0.0, 1.0, 1.0]] * 2))


def test_no_cpu_fallback_in_product(pkg):
    """the product package must not import or reference the oracle"""
    pkg_dir = ROOT / "volumefluxes.jl_b200"
    for p in list(pkg_dir.glob("*.py")) + list((pkg_dir / "csrc").glob("*")):
        if p.suffix in (".py", ".cu", ".cuh", ".h"):
            txt = p.read_text()
            assert "liborc" not in txt and "swe_oracle" not in txt and "import oracle" not in txt, p


def test_synthetic_terrain_is_deterministic():
    syn = cases.synthetic()
    a = syn.terrain_corners(64, 48, seed=1234, wavelength=32.0)
    b = syn.terrain_corners(64, 48, seed=1234, wavelength=32.0)
    assert a.shape == (53, 69) and a.dtype == np.float32 and np.array_equal(a, b)
    assert np.ptp(a) > 0.5 and np.abs(np.diff(a, axis=1)).max() < 0.5
    c = syn.cell_average(a)
    assert c.shape == (52, 68)


def test_cost_balanced_partition(pkg):
    """multigpu.partition_rows: equal heights without a cost; cost-equalising cuts with one (the dt all-gather
    makes every rank wait for the slowest strip, so strips are cut by estimated work, not by rows)"""
    mg = pkg.multigpu
    assert mg.partition_rows(83, 3) == ([0, 28, 56], [28, 28, 27])
    wet = np.r_[np.zeros(40), np.ones(43)]
    offs, sizes = mg.partition_rows(83, 3, mg.row_cost_from_wet_fraction(wet, dry_cost=0.5))
    assert offs[0] == 0 and sum(sizes) == 83 and sizes[0] > sizes[2]          # dry rows are cheaper -> taller strip
    cost = mg.row_cost_from_wet_fraction(wet, dry_cost=0.5)
    work = [cost[o:o + n].sum() for o, n in zip(offs, sizes)]
    assert max(work) - min(work) <= 2 * cost.max()
    assert mg.partition_rows(16, 4, np.r_[np.zeros(12), np.ones(4)])[1] == [4, 4, 4, 4]   # never below 4 rows
    with pytest.raises(pkg.ArgumentError):
        mg.partition_rows(83, 3, np.ones(82))


def test_rebalance_from_measured_times(pkg):
    """second pass of the strip balancing (bench.py): a rank that measured 20 % slower than the model predicted gets
    fewer rows, and re-measuring with a cost that is exact leaves the cuts alone"""
    mg = pkg.multigpu
    ny, P = 4000, 4
    model = np.ones(ny)
    offs, sizes = mg.partition_rows(ny, P, model)
    assert sizes == [1000] * 4
    true_speed = np.array([1.0, 1.2, 1.0, 0.9])                       # per-row cost of each rank's GPU
    measured = [sizes[r] * true_speed[r] for r in range(P)]
    cost2 = mg.rescale_cost_by_measured(model, offs, sizes, measured)
    offs2, sizes2 = mg.partition_rows(ny, P, cost2)
    assert sum(sizes2) == ny and sizes2[1] < 1000 < sizes2[3]
    # rows kept their rank's measured scale, so the predicted per-rank time is flat to within the moved rows
    t2 = [cost2[o:o + n].sum() for o, n in zip(offs2, sizes2)]
    assert max(t2) / min(t2) < 1.002
    with pytest.raises(pkg.ArgumentError):
        mg.rescale_cost_by_measured(model, offs, sizes, [1.0, 1.0, 0.0, 1.0])


def test_x2_kernels_sass_has_packed_fp32_and_async_copies():
    """Build-quality guard (no GPU needed): the shipped Float32 stage kernels really are the sm_100a design DESIGN.md
    describes -- packed FFMA2/FMUL2/FADD2 arithmetic, LDGSTS (cp.async) ring fills, no barrier -- and ptxas reports no
    spills for the two instantiations the bench runs (stage 1 at 128 registers, stage 2 at 168)."""
    import shutil
    obj = ROOT / "volumefluxes.jl_b200" / "csrc" / "engine_fast.o"
    log = ROOT / "volumefluxes.jl_b200" / "csrc" / "engine_fast.ptxas.log"
    if shutil.which("cuobjdump") is None or not obj.exists() or not log.exists():
        pytest.skip("needs the CUDA toolkit and a built engine_fast.o")
    sass = subprocess.run(["cuobjdump", "-sass", str(obj)], capture_output=True, text=True, check=True).stdout
    kernels = {}
    name = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = []
        elif name and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            kernels[name].append(line)
    # (the kernels live in a per-translation-unit inline namespace, tu_namespace.h: match on the template arguments)
    def find(args):
        hits = [k for k in kernels if "k_stage_x2" in k and args in k]
        assert len(hits) == 1, f"k_stage_x2<{args}> not (uniquely) in engine_fast.o: {hits}"
        return hits[0]
    k1 = find("ILi1ELb1ELi1ELi0ELb0EE")     # <RECON=1, HASB=true, STAGE=1, FRIC=0, FACTOR=false>
    k2 = find("ILi1ELb1ELi2ELi1ELb0EE")     # <RECON=1, HASB=true, STAGE=2, FRIC=1, FACTOR=false>
    for k in (k1, k2):
        assert k in kernels, f"{k} not in engine_fast.o"
        body = "\n".join(kernels[k])
        for op in ("FFMA2", "FMUL2", "FADD2", "LDGSTS", "MUFU.RCP", "SHFL"):
            assert op in body, f"{op} missing from {k}"
        assert "BAR.SYNC" not in body, "the x2 kernels are barrier-free by design"
        assert body.count("FFMA2") + body.count("FMUL2") + body.count("FADD2") > 800      # packed ops carry the arithmetic
    txt = log.read_text()
    for k, regs in ((k1, 128), (k2, 168)):
        m = re.search(re.escape(k) + r"'.*?\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", txt)
        assert m, f"no ptxas record for {k}"
        assert (int(m.group(2)), int(m.group(3))) == (0, 0), f"{k} spills: {m.group(0)}"
        assert int(m.group(4)) <= regs


def test_partition_rows_properties(pkg):
    """property test (hypothesis) of the strip partition: exact cover, >= 4 rows each, and with a cost the heaviest
    strip is within one row's cost (plus the 4-row floor) of the ideal share"""
    from hypothesis import given, settings, strategies as st
    mg = pkg.multigpu

    @settings(max_examples=200, deadline=None)
    @given(st.integers(1, 8), st.integers(0, 400), st.integers(0, 2 ** 31 - 1), st.booleans())
    def check(nranks, extra, seed, with_cost):
        ny = 4 * nranks + extra
        cost = None
        if with_cost:
            rng = np.random.default_rng(seed)
            cost = mg.row_cost_from_wet_fraction(rng.random(ny) ** rng.integers(1, 4), dry_cost=float(rng.uniform(0.2, 1.0)))
        offs, sizes = mg.partition_rows(ny, nranks, cost)
        assert len(offs) == len(sizes) == nranks and sum(sizes) == ny and offs[0] == 0
        assert all(s >= 4 for s in sizes)
        assert all(offs[r + 1] == offs[r] + sizes[r] for r in range(nranks - 1))
        if cost is None:
            assert max(sizes) - min(sizes) <= 1
        elif ny >= 16 * nranks:
            loads = [float(cost[o:o + s].sum()) for o, s in zip(offs, sizes)]
            assert max(loads) <= cost.sum() / nranks + 2 * cost.max() + 1e-9

    check()


def test_grid_helpers_match_reference_tests(pkg):
    """test_grid.jl:24-48 and test_grid_2d.jl:25-74 on the host mirror (CartesianGrid 1D / 2D, cell_faces,
    cell_centers incl. the ghost region, compute_dx / compute_dy)."""
    P = pkg
    nx, ny, gc = 10, 6, 2
    g1 = P.CartesianGrid(nx, gc=2, extent=[0, nx])
    assert g1.dim == 1 and g1.totalcells == (nx + 4,) and P.compute_dx(g1) == 1.0
    g = P.CartesianGrid(nx, ny, gc=gc, extent=[[0, nx], [0, ny]])
    assert P.compute_dx(g) == 1.0 and P.compute_dy(g) == 1.0 and P.interior_size(g) == (nx, ny)
    xf, yf = np.arange(nx + 1.0), np.arange(ny + 1.0)
    xc, yc = np.arange(0.5, nx), np.arange(0.5, ny)
    assert np.array_equal(P.cell_faces(g, P.XDIR), xf) and np.array_equal(P.cell_faces(g, P.YDIR), yf)
    assert np.array_equal(P.cell_centers(g, P.XDIR), xc) and np.array_equal(P.cell_centers(g, P.YDIR), yc)
    f2, c2 = P.cell_faces(g), P.cell_centers(g)              # [j, i, (x, y)] == the reference's [i, j][1:2]
    assert f2.shape == (ny + 1, nx + 1, 2) and c2.shape == (ny, nx, 2)
    assert np.array_equal(f2[..., 0], np.tile(xf, (ny + 1, 1))) and np.array_equal(f2[..., 1], np.tile(yf[:, None], (1, nx + 1)))
    assert np.array_equal(c2[..., 0], np.tile(xc, (ny, 1))) and np.array_equal(c2[..., 1], np.tile(yc[:, None], (1, nx)))
    fg = P.cell_faces(g, interior=False)
    assert np.allclose(fg[2:-2, 2:-2], f2, atol=1e-14)
    cg = P.cell_centers(g, interior=False)
    assert np.allclose(cg[2:-2, 2:-2], c2, atol=1e-14)
    for j in (1, 2):
        for i in (1, 2):
            assert np.allclose(fg[j - 1, i - 1], [-3 + i, -3 + j], atol=1e-14)
            assert np.allclose(fg[-j, -i], [nx + (gc - i + 1), ny + (gc - j + 1)], atol=1e-14)
            assert np.allclose(cg[j - 1, i - 1], [-3 + i + 0.5, -3 + j + 0.5], atol=1e-14)
            assert np.allclose(cg[-j, -i], [nx + (gc - i) + 0.5, ny + (gc - j) + 0.5], atol=1e-14)
    # 1D, ghost region (test_grid.jl:40-48)
    syn = P.synthetic
    xfg = syn.cell_faces(nx, 2, 0.0, float(nx), interior=False)
    assert np.allclose(xfg[2:-2], xf, atol=1e-14) and np.allclose(xfg[:2], [-2, -1], atol=1e-14) and np.allclose(xfg[-2:], [11, 12], atol=1e-14)
    xcg = syn.cell_centers(nx, 2, 0.0, float(nx), interior=False)
    assert np.allclose(xcg[:2], [-1.5, -0.5], atol=1e-14) and np.allclose(xcg[-2:], [10.5, 11.5], atol=1e-14)
    # inner_cells (grid.jl:213-220; test_for_each_inner_cell_2d.jl)
    assert P.inner_cells(g, P.XDIR) == (nx, ny) and P.inner_cells(g, P.XDIR, 1) == (nx + 2, ny)
    assert P.inner_cells(g, P.YDIR, 1) == (nx, ny + 2)
