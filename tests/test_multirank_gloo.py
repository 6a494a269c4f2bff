"""world_size = 2 (and 3) tests of the N > 1 host logic on CPU with the gloo backend: the strip
decomposition used by the multi-GPU path (partition_rows / strip_case, physical-side flags, periodic
wrap-around, global dt from all-reduced wave speeds, 2-row halo trades) reproduces the single-domain
oracle BITWISE.  No GPU is involved: compute is the CPU oracle, driven by tests/strip_oracle.py."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, dtype, nsteps, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    import torch.distributed as dist
    import cases
    import strip_oracle
    from oracle import oracle
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P = cases.load_package()
    case, state = CASES[name](dtype)
    so = strip_oracle.StripOracle(oracle, P.multigpu, case, state, rank, world, strip_oracle.TorchComm())
    mts = case.get("maximum_timestep") or float("inf")
    dts = []
    for _ in range(nsteps):
        so.step(mts)
        dts.append(float(so.dt))
    np.savez(Path(out_dir) / f"rank{rank}.npz", state=np.stack(so.local_interior()), dts=np.array(dts), t=float(so.t),
             y0=so.y0)
    dist.barrier()
    dist.destroy_process_group()


def _cases():
    import cases
    return {
        "lake_wall_rk2": lambda dt: cases.lake_case(n=48, ny=41, dtype=dt, level=0.0, wavelength=24.0),
        "dry_periodic_rk2": lambda dt: cases.dry_states_2d(dtype=dt, n=32),
        "terrain_horton_fe": lambda dt: cases.terrain_case(n=40, dtype=dt, level=0.0, wavelength=24.0,
                                                           friction="bh2021", stepper="fe"),
    }


sys.path.insert(0, str(ROOT / "tests"))
CASES = _cases()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_strips_reproduce_single_domain_oracle(tmp_path, oracle, name, dtype, world):
    import torch.multiprocessing as mp
    nsteps = 12
    port = _free_port()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, dtype, nsteps, str(tmp_path))) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    case, state = CASES[name](dtype)
    o = oracle.OracleSim(case)
    o.set_state_interior(*state)
    mts = case.get("maximum_timestep") or float("inf")
    dts = []
    for _ in range(nsteps):
        o.step(mts)
        dts.append(o.dt)
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    got = np.concatenate([p["state"] for p in parts], axis=1)
    assert np.array_equal(got, np.stack(o.get_state_interior()))
    for p in parts:
        assert np.array_equal(p["dts"], np.array(dts)) and float(p["t"]) == o.t
    assert [int(p["y0"]) for p in parts] == sorted(int(p["y0"]) for p in parts)
