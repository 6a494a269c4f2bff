"""y-strip domain decomposition of one Simulator over the GPUs of a box.

The reference is single-device; this is the new capability BASELINE.json's config 5 asks for.
One process per GPU (torchrun): every rank owns ny/P interior rows plus 2 ghost rows per side; the
x-direction BC is local on every rank, the y-direction BC only on the first / last rank (periodic-y
becomes an exchange between rank 0 and rank P-1; NeumannBC's low-y image a one-directional push from rank P-1 to
rank 0, with the far end's bathymetry rows handed over once: swe_b200_set_wrap_bathymetry).  The data path never touches the host: the stage
kernels' neighbours push halo rows and CFL maxima into each other's memory over NVLink
(csrc/kernels_exchange.cuh); torch.distributed is used ONLY to all-gather the 128-byte window
descriptors once at start-up (and by tests / bench to gather results).
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, List, Sequence

import numpy as np

from . import _lib
from . import api

IPC_BYTES = 128


def partition_rows(ny: int, nranks: int, row_cost=None):
    """Rows [offset, offset+size) of the global interior owned by each rank.

    Without `row_cost` the sizes differ by at most 1.  With `row_cost` (one non-negative number per interior
    row, e.g. `row_cost_from_wet_fraction`) the cuts equalise the summed cost instead: every rank waits for the
    slowest one at the per-step CFL all-gather, and wet rows cost ~1.7x dry ones (the flux evaluation is skipped
    for warps whose faces are all dry), so equal heights are not equal work."""
    if nranks < 1 or ny < 4 * nranks:
        raise api.ArgumentError("every strip needs at least 4 interior rows")
    if row_cost is None:
        base, rem = divmod(ny, nranks)
        sizes = [base + (1 if r < rem else 0) for r in range(nranks)]
    else:
        c = np.asarray(row_cost, dtype=np.float64)
        if c.shape != (ny,) or np.any(c < 0) or not np.all(np.isfinite(c)):
            raise api.ArgumentError("row_cost must hold one finite non-negative number per interior row")
        cum = np.concatenate([[0.0], np.cumsum(c + 1e-12)])
        cuts = [0]
        for r in range(1, nranks):
            k = int(np.searchsorted(cum, cum[-1] * r / nranks))
            k = min(max(k, cuts[-1] + 4), ny - 4 * (nranks - r))     # at least 4 rows for everybody
            cuts.append(k)
        cuts.append(ny)
        sizes = [cuts[r + 1] - cuts[r] for r in range(nranks)]
    offsets = [sum(sizes[:r]) for r in range(nranks)]
    return offsets, sizes


def row_cost_from_wet_fraction(wet_fraction_per_row, dry_cost: float = 0.6):
    """Cost model of one interior row for `partition_rows`: dry cells cost `dry_cost` of a wet cell."""
    w = np.clip(np.asarray(wet_fraction_per_row, dtype=np.float64), 0.0, 1.0)
    return dry_cost + (1.0 - dry_cost) * w


def rescale_cost_by_measured(row_cost, offsets, sizes, measured):
    """Second pass of the strip balancing: `measured[r]` is the time rank r actually spent in its stage kernels on the
    strip [offsets[r], offsets[r] + sizes[r]) cut with `row_cost`.  The model is rescaled strip by strip so that its
    sum over each strip equals the measurement (GPU-to-GPU clock / power differences, item-schedule effects and
    whatever else the model misses end up in the scale), and `partition_rows` can cut again with the result."""
    c = np.asarray(row_cost, dtype=np.float64)
    m = np.asarray(measured, dtype=np.float64)
    if len(offsets) != len(sizes) or len(m) != len(sizes) or sum(sizes) != len(c) or np.any(m <= 0):
        raise api.ArgumentError("rescale_cost_by_measured: one positive measurement per strip, strips covering all rows")
    out = np.empty_like(c)
    for o, n, tm in zip(offsets, sizes, m):
        s = c[o:o + n].sum()
        out[o:o + n] = c[o:o + n] * (tm / s if s > 0 else 1.0)
    return out


def strip_case(case: dict, state: Sequence[np.ndarray], rank: int, nranks: int, row_cost=None):
    """The sub-problem of `rank`: same parameters, its rows of the state, the bathymetry corner rows
    (owned rows + 2 ghost rows per side + 1) and of the infiltration factor plane."""
    gc = case.get("gc", 2)
    ny = case["ny"]
    offsets, sizes = partition_rows(ny, nranks, row_cost)
    y0, n = offsets[rank], sizes[rank]
    sub = dict(case)
    sub["ny"] = n
    B = case.get("B")
    if B is not None and not np.isscalar(B):
        Bg = np.asarray(B)
        sub["B"] = np.ascontiguousarray(Bg[y0:y0 + n + 2 * gc + 1])
        # ghost rows that are a wrap-around image of the far end of the domain (periodic: both ends; NeumannBC: the low
        # ghosts are the image of the top interior rows, bc.jl:56): stage 2 needs the bathymetry of the SOURCE cells
        bc = case.get("bc", "periodic")
        if nranks > 1 and rank == 0 and bc in ("periodic", "neumann"):
            sub["wrap_B_lo"] = np.ascontiguousarray(Bg[ny:ny + gc + 1])              # corner rows of global cell rows ny .. ny+gc-1
        if nranks > 1 and rank == nranks - 1 and bc == "periodic":
            sub["wrap_B_hi"] = np.ascontiguousarray(Bg[gc:2 * gc + 1])              # corner rows of global cell rows gc .. 2gc-1
    infil = case.get("infiltration")
    if isinstance(infil, dict) and not np.isscalar(infil.get("factor", 1.0)):
        f = dict(infil)
        f["factor"] = np.ascontiguousarray(np.asarray(infil["factor"])[y0:y0 + n + 2 * gc])
        sub["infiltration"] = f
    sub_state = tuple(np.ascontiguousarray(np.asarray(a)[y0:y0 + n]) for a in state)
    return sub, sub_state, y0, n


def _blob(sim: api.Simulator) -> bytes:
    buf = (C.c_ubyte * IPC_BYTES)()
    sim._check(_lib.lib().swe_b200_get_ipc_handle(sim._h, buf))
    return bytes(buf)


def _open(sim: api.Simulator, blobs: List[bytes]):
    raw = b"".join(blobs)
    buf = (C.c_ubyte * len(raw)).from_buffer_copy(raw)
    sim._check(_lib.lib().swe_b200_open_peers(sim._h, buf))


class StripSimulator:
    """This rank's strip of a global Simulator (one process per GPU).  `all_gather(bytes) -> [bytes]*P`
    is any host-side all-gather (see `torch_all_gather`).  Build it from the global problem
    (`from_global`) or from an already cut strip (constructor)."""

    def __init__(self, sub_case: dict, sub_state, rank: int, nranks: int, global_ny: int, y_offset: int,
                 all_gather: Callable[[bytes], List[bytes]], device: int = 0, strict: bool = False):
        self.rank, self.nranks = rank, nranks
        self._all_gather = all_gather
        self.y0, self.ny, self.global_ny = y_offset, sub_case["ny"], global_ny
        self.sim = api.simulator_from_case(sub_case, strict=strict, fused=True, device=device, rank=rank,
                                           nranks=nranks, global_ny=global_ny, y_offset=y_offset)
        if nranks > 1:
            _open(self.sim, all_gather(_blob(self.sim)))
        self.sim.set_state_interior(*sub_state)      # collective: pushes the first halo rows to the neighbours

    @classmethod
    def from_global(cls, case: dict, state, rank: int, nranks: int, all_gather, device: int = 0, strict: bool = False):
        sub, sub_state, y0, _ = strip_case(case, state, rank, nranks)
        return cls(sub, sub_state, rank, nranks, case["ny"], y0, all_gather, device=device, strict=strict)

    def close(self):
        """Collective tear-down: every rank unmaps its peers' windows, then all ranks meet (the all-gather doubles as the
        barrier), then the handle is destroyed -- freeing a window a peer still has mapped is undefined in CUDA IPC."""
        if self.sim is None:
            return
        if self.nranks > 1:
            self.sim._check(_lib.lib().swe_b200_close_peers(self.sim._h))
            self._all_gather(b"closed")
        self.sim.close()
        self.sim = None

    def run_steps(self, n, maximum_timestep=None):
        self.sim.run_steps(n, maximum_timestep)

    def run_steps_timed(self, n, maximum_timestep=None):
        return self.sim.run_steps_timed(n, maximum_timestep)

    def simulate_to_time(self, T, match_endtime=True, maximum_timestep=None):
        return api.simulate_to_time(self.sim, T, match_endtime=match_endtime, maximum_timestep=maximum_timestep)

    def local_interior(self):
        return self.sim.get_state_interior()


def torch_all_gather(payload: bytes) -> List[bytes]:
    """all-gather of a small byte string over the default torch.distributed process group"""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    mine = torch.tensor(list(payload), dtype=torch.uint8, device=dev)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)
    return [bytes(t.cpu().tolist()) for t in out]


class LocalStripGroup:
    """All strips inside ONE process, one per visible device (used by tests and single-process runs).
    Kernels of different strips wait for each other on the GPUs, so only the asynchronous entry points
    (run_steps) may be used before `sync()`."""

    def __init__(self, case: dict, state, devices: Sequence[int], strict: bool = False, row_cost=None):
        P = len(devices)
        self.case = case
        self.sims, self.rows = [], []
        subs = []
        for r, dev in enumerate(devices):
            sub, sub_state, y0, n = strip_case(case, state, r, P, row_cost)
            self.sims.append(api.simulator_from_case(sub, strict=strict, fused=True, device=dev, rank=r, nranks=P,
                                                     global_ny=case["ny"], y_offset=y0))
            subs.append(sub_state)
            self.rows.append((y0, n))
        if P > 1:
            blobs = [_blob(s) for s in self.sims]
            for s in self.sims:
                _open(s, blobs)
        for s, st in zip(self.sims, subs):
            s.set_state_interior(*st)

    def run_steps(self, n, maximum_timestep=None):
        # interleave so that no rank's launch queue runs far ahead of its neighbours
        for _ in range(int(n)):
            for s in self.sims:
                s.run_steps(1, maximum_timestep)

    def sync(self):
        for s in self.sims:
            s.sync()

    def gather_interior(self):
        self.sync()
        parts = [s.get_state_interior() for s in self.sims]
        return [np.concatenate([p[k] for p in parts], axis=0) for k in range(3)]

    @property
    def t(self):
        return self.sims[0].t

    @property
    def dt(self):
        return self.sims[0].dt
