"""Host-side driver of a y-strip decomposition on the CPU oracle (test infrastructure).

Every rank runs the oracle on its strip (cut with the product's volumefluxes.jl_b200.multigpu.strip_case),
computes dt from the all-reduced wave speeds exactly like timestep_computer (simulator.jl:147-150) and
trades the 2 outermost owned rows with its y-neighbours after every substep through `comm`.
Used by the gloo world_size=2 tests: validates partitioning, strip slicing of B / factor, which sides are
physical boundaries, periodic wrap-around neighbours and the global-dt protocol -- without a GPU.
"""
import math

import numpy as np


class Comm:
    """minimal interface: allreduce_max(np.ndarray) and sendrecv(dst, send_array, src, recv_shape, dtype)"""

    def allreduce_max(self, a):
        raise NotImplementedError

    def exchange(self, lo, hi, to_lo, to_hi):
        """send `to_lo` to rank lo and `to_hi` to rank hi (None = no neighbour); returns (from_lo, from_hi)"""
        raise NotImplementedError


class TorchComm(Comm):
    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist

    def allreduce_max(self, a):
        t = self.torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64).copy())
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.numpy()

    def exchange(self, lo, hi, to_lo, to_hi):
        torch, dist = self.torch, self.dist
        reqs, out = [], [None, None]
        bufs = []
        for idx, (peer, arr) in enumerate(((lo, to_lo), (hi, to_hi))):
            if peer is None:
                continue
            recv = torch.empty(arr.shape, dtype=torch.from_numpy(arr).dtype)
            send = torch.from_numpy(np.ascontiguousarray(arr).copy())
            bufs.append((idx, recv))
            # tag by direction so that a 2-rank periodic ring (lo == hi) does not mix the two messages
            reqs.append(dist.isend(send, peer, tag=idx))
            reqs.append(dist.irecv(recv, peer, tag=1 - idx))
        for r in reqs:
            r.wait()
        for idx, recv in bufs:
            out[idx] = recv.numpy()
        return out[0], out[1]


class StripOracle:
    def __init__(self, oracle_mod, multigpu, case, state, rank, nranks, comm):
        self.case, self.rank, self.P, self.comm = case, rank, nranks, comm
        sub, sub_state, self.y0, self.ny = multigpu.strip_case(case, state, rank, nranks)
        # dy must be the GLOBAL cell height: keep the strip's extent proportional to its rows
        ext = list(case.get("extent", [0.0, 1.0, 0.0, 1.0]))
        dy = (ext[3] - ext[2]) / case["ny"]
        sub["extent"] = [ext[0], ext[1], ext[2] + self.y0 * dy, ext[2] + self.y0 * dy + self.ny * dy]
        self.dx = (ext[1] - ext[0]) / case["nx"]
        self.dy = dy
        self.o = oracle_mod.OracleSim(sub)
        self.dtype = self.o.np_dtype
        periodic = case.get("bc", "periodic") == "periodic"
        self.lo = rank - 1 if rank > 0 else (nranks - 1 if periodic and nranks > 1 else None)
        self.hi = rank + 1 if rank < nranks - 1 else (0 if periodic and nranks > 1 else None)
        self.ylo_phys = self.lo is None
        self.yhi_phys = self.hi is None
        self.gc = case.get("gc", 2)
        self.t = self.dtype.type(case.get("t0", 0.0))
        self.dt = self.dtype.type(0)
        self.cfl = self.dtype.type(case.get("cfl", 0.25))
        self.nsub = 2 if case.get("stepper", "rk2") == "rk2" else 1
        self.o.set_state_interior(*sub_state)      # interior + BC on all sides ...
        self._halo(0)                              # ... then the neighbours' rows replace the inner y ghosts

    def _halo(self, which):
        g = self.gc
        full = self.o.get_state_full(which)
        to_lo = full[:, g:2 * g, :] if self.lo is not None else None
        to_hi = full[:, -2 * g:-g, :] if self.hi is not None else None
        from_lo, from_hi = self.comm.exchange(self.lo, self.hi, to_lo, to_hi)
        if from_lo is not None:
            full[:, :g, :] = from_lo
        if from_hi is not None:
            full[:, -g:, :] = from_hi
        self.o.set_state_full(full, which)

    def step(self, max_dt=math.inf):
        o = self.o
        for s in range(self.nsub):
            out = s + 1
            o.zero(out)
            o.set_time(float(self.t))
            ms = o.add_time_derivative(s, out, float(self.t))
            if s == 0:
                ms = self.comm.allreduce_max(np.array(ms, dtype=np.float64))
                ax, ay = self.dtype.type(ms[0]), self.dtype.type(ms[1])
                with np.errstate(divide="ignore"):
                    ddt = min(np.float64(self.dx) / np.float64(ax), np.float64(self.dy) / np.float64(ay))
                dt = min(float(np.float64(self.cfl) * ddt), max_dt)
                o.axpy(out, s, dt, rounded=False)
                self.dt = self.dtype.type(dt)
            else:
                o.axpy(out, s, float(self.dt), rounded=True)
                o.rk_average(out, 0)
            o.implicit_friction(out, s, float(self.dt))
            o.post_proc(out)
            o.update_bc_sides(out, True, True, self.ylo_phys, self.yhi_phys)
            self._halo(out)
        o.rotate()
        self.t = self.dtype.type(self.t + self.dt)

    def local_interior(self):
        return self.o.get_state_interior()
