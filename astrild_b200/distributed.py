"""Slab-decomposed P(k) across the GPUs of one node: one process per GPU, torch.distributed/NCCL.

astrild itself is single-rank; the MPI decomposition lives unused inside
pmesh/pfft.  This module is the B200 counterpart of that decomposition for the same path
(/root/reference/src/astrild/particles/hutils/stats_subfind.py:125-150):

  1. route      every particle to the rank owning the x-slab of floor(g_x)   (all-to-all-v)
  2. deposit    into [1 ghost | n0 owned | 2 ghost] planes                    (local kernel)
  3. ghosts     ghost planes -> ring neighbours, added to their owned planes  (send/recv)
  4. 2-D r2c    over (y, z) on the owned planes                               (cuFFT, local)
  5. transpose  x <-> y: rank s receives y in [s N/P, (s+1) N/P) for all x     (all-to-all)
  6. 1-D c2c    along x                                                        (cuFFT, local)
  7. binning    fused kernel on the transposed [x][y_local][z] grid            (local kernel)
  8. reduce     shell sums, mode counts and total mass                         (all-reduce)

Nothing is transposed back: the binning kernel only needs per-axis k tables (exactly what
pfft's PFFT_TRANSPOSED_OUT gives nbodykit).  Mode counts are integers, so they are identical to
the single-GPU result; the float64 sums differ only in summation order.

The compute stages go through a *backend* object; the product backend is CudaSlabBackend
(libastrild_pk.so).  tests/ inject a NumPy backend to exercise this file's exchange logic with the
gloo backend on CPUs -- that backend lives in tests/, never here.
"""
from __future__ import annotations

import ctypes as ct
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, tables
from ._lib import AstrildPkError
from .engine import PkEngine, _ptr


class CudaSlabBackend:
    """Compute stages of one rank on its GPU (libastrild_pk.so)."""

    def __init__(self, N: int, L: float, x0: int, n0: int, nranks: int, device):
        self.eng = PkEngine(N, L, device, x0=x0, n0=n0)
        self.device = self.eng.device
        self.N, self.L, self.n0, self.nranks = N, L, n0, nranks
        self._counts = torch.zeros(2 * nranks, dtype=torch.int64, device=self.device)
        # high priority: its kernels (exchange, first mesh's ghosts / FFT / transpose) take SMs as deposit CTAs retire
        self.side_stream = torch.cuda.Stream(self.device, priority=-1)

    # -- 1. routing -----------------------------------------------------------------------
    def route(self, pos, mass, pos_scale: float):
        """Particles of this rank that belong to another slab, grouped by destination:
        (positions (n,3), masses or None, per-destination counts).  Everything else stays put."""
        return self.route_end(self.route_begin(pos, mass, pos_scale))

    def route_begin(self, pos, mass, pos_scale: float, capacity: int | None = None, counts=None) -> dict:
        """Launches the routing kernels on the current stream and returns without waiting for them.
        counts: int64[2 * nranks] device tensor that receives the per-destination counts (default: the backend's own)."""
        eng = self.eng
        p0, p1, p2, layout, dt, npart, keep = eng._positions(pos)
        m = None
        if mass is not None and not np.isscalar(mass):          # a scalar weight is the caller's to apply (SlabPk.power)
            m = eng._to_device(mass).to(dt).contiguous()
        code = _lib.APK_F32 if dt == torch.float32 else _lib.APK_F64
        if capacity is None:
            # what left last time (+ 25 %) if that is known: unrouted input overflows npart / 8 every time, and an
            # overflow costs a second pass after a device synchronise
            capacity = min(max(npart, 1), max(1 << 16, npart // 8, int(1.25 * getattr(self, "_last_leavers", 0))))
        # the leavers are staged in the plan workspace: capacity * (3 + [mass]) * itemsize + 64 bytes; the workspace holds
        # at least 12 bytes per particle it is sized for
        item = 4 if dt == torch.float32 else 8
        stage_rows = (capacity * item * (4 if m is not None else 3) + 64 + 11) // 12
        eng.ensure_workspace(max(npart, stage_rows), m is not None)
        out_pos = torch.empty((capacity, 3), dtype=dt, device=self.device)
        out_mass = torch.empty(capacity, dtype=dt, device=self.device) if m is not None else None
        cnt = self._counts if counts is None else counts
        _lib.call("apk_route_particles", eng._plan, _ptr(p0), _ptr(p1), _ptr(p2), layout, code, float(pos_scale),
                  _ptr(m), code, int(npart), self.nranks, _ptr(cnt), int(capacity), _ptr(out_pos),
                  _ptr(out_mass), eng.stream)
        return {"pos": pos, "mass": mass, "pos_scale": pos_scale, "capacity": capacity, "out_pos": out_pos,
                "out_mass": out_mass, "keep": (keep, m), "counts": cnt}

    def route_end(self, h: dict):
        """Reads the per-destination counts (blocks the host on the CURRENT stream, which must be ordered after
        route_begin's) and returns (positions, masses, counts).  More leavers than the staging buffer holds (more
        than 1/8 of the particles change slab: rare) means a second, larger pass after a device synchronise."""
        counts = h["counts"][: self.nranks].cpu().tolist()
        total = int(sum(counts))
        self._last_leavers = total
        if total > h["capacity"]:
            torch.cuda.synchronize(self.device)       # nothing else may be using the plan workspace
            return self.route_end(self.route_begin(h["pos"], h["mass"], h["pos_scale"], capacity=total))
        om = h["out_mass"]
        return h["out_pos"][:total], (None if om is None else om[:total]), counts

    def empty_like_rows(self, like: torch.Tensor, rows: int) -> torch.Tensor:
        return torch.empty((rows,) + tuple(like.shape[1:]), dtype=like.dtype, device=self.device)

    # -- 2./3. deposit and ghost planes ---------------------------------------------------------
    def deposit(self, pos_aos, mass, resampler: str, shift: float, pos_scale: float, out=None, method: str = "auto"):
        """Deposit into the slab buffer (with ghost planes); out != None accumulates into it."""
        return self.eng.deposit(pos_aos, mass, resampler, shift, pos_scale, method, out=out, zero=out is None)

    def deposit_pair(self, pos_aos, mass, resampler: str, pos_scale: float, out=None, method: str = "auto",
                     timed: bool = True):
        """Both interlaced twins (shift 0, 0.5) from one partition; out = (mesh, mesh_shifted) accumulates.
        timed=False keeps this call out of the plan's per-kernel timing (it runs beside another deposit)."""
        was = self.eng.timing
        if not timed and was:
            self.eng.enable_timing(False)
        try:
            return self.eng.deposit_pair(pos_aos, mass, resampler, pos_scale, method, out=out, zero=out is None)
        finally:
            if not timed and was:
                self.eng.enable_timing(True)

    def prepare_ffts(self, ny: int) -> None:
        """Create both slab FFT plans now, so that the next workspace request already includes their work areas."""
        _lib.call("apk_plan_prepare_fft2d", self.eng._plan)
        _lib.call("apk_plan_prepare_fft1d", self.eng._plan, int(ny))

    def set_first_mesh_event(self, event) -> None:
        """event: a recorded torch.cuda.Event or None; see apk_plan_set_first_mesh_event."""
        _lib.call("apk_plan_set_first_mesh_event", self.eng._plan,
                  ct.c_void_p(event.cuda_event) if event is not None else None)

    def zeroed_meshes(self, count: int) -> list:
        """Slab buffers (with ghost planes), zeroed on the current stream."""
        return [self.eng.new_mesh(ghosts=True).zero_() for _ in range(count)]

    def accumulate(self, dst: torch.Tensor, src: torch.Tensor) -> None:
        assert dst.is_contiguous() and src.is_contiguous() and dst.numel() == src.numel()
        _lib.call("apk_mesh_accumulate", self.eng._plan, _ptr(dst), _ptr(src), dst.numel(), self.eng.stream)

    def mesh_sum(self, owned: torch.Tensor) -> torch.Tensor:
        out = torch.zeros(1, dtype=torch.float64, device=self.device)
        _lib.call("apk_padded_mesh_sum", self.eng._plan, _ptr(owned), _ptr(out), self.eng.stream)
        return out

    @staticmethod
    def dc_sum(grid2d: torch.Tensor) -> torch.Tensor:
        """Sum of the owned planes read off the (ky, kz) = (0, 0) column of their 2-D transforms (cuFFT is
        un-normalised): n0 values instead of a pass over the mesh."""
        return grid2d[:, 0, 0].real.sum(dtype=torch.float64).reshape(1)

    # -- 4./6. FFT stages ---------------------------------------------------------------------------
    def fft2d(self, owned: torch.Tensor) -> torch.Tensor:
        eng = self.eng
        eng.ensure_workspace(0, False)
        _lib.call("apk_fft_r2c_2d", eng._plan, _ptr(owned), eng.stream)
        return torch.view_as_complex(owned.view(self.n0, self.N, eng.Nk, 2))

    def fft1d(self, grid: torch.Tensor, ny: int) -> torch.Tensor:
        eng = self.eng
        _lib.call("apk_plan_prepare_fft1d", eng._plan, int(ny))     # plan first: it sizes the workspace
        eng.ensure_workspace(0, False)
        _lib.call("apk_fft_c2c_1d", eng._plan, _ptr(grid), int(ny), eng.stream)
        return grid

    # -- 5. transpose over NVLink peer memory -------------------------------------------------------
    def setup_p2p(self, group, nfields: int) -> bool:
        """Peer-mapped receive buffers (torch symmetric memory) for the fused pack + peer-store transpose.
        Returns False (and the caller uses the NCCL all-to-all) if symmetric memory is unavailable."""
        if getattr(self, "_p2p_fields", 0) >= nfields:
            return self._p2p is not None
        self._p2p, self._p2p_fields = None, nfields
        try:
            import torch.distributed._symmetric_memory as symm_mem
            ny, Nk = self.N // self.nranks, self.N // 2 + 1
            buf = symm_mem.empty((nfields, self.N, ny, Nk), dtype=torch.complex64, device=self.device)
            name = (group if group is not None else dist.group.WORLD).group_name
            hdl = symm_mem.rendezvous(buf, name)
            self._p2p = (buf, hdl)
        except Exception as e:  # noqa: BLE001 -- any failure: keep the NCCL path
            self._p2p_error = repr(e)
            self._p2p = None
        return self._p2p is not None

    def transpose_p2p(self, grids: list) -> list:
        buf, hdl = self._p2p
        hdl.barrier(channel=0)                    # every rank is done reading the previous step's buffers
        for f, g in enumerate(grids):
            self.transpose_p2p_store(f, g)
        hdl.barrier(channel=1)                    # every rank's blocks have landed
        return [buf[f] for f in range(len(grids))]

    def transpose_p2p_store(self, f: int, g: torch.Tensor, stream=None) -> torch.Tensor:
        """Pack + peer-store field f into every rank's receive buffer (no synchronisation: see p2p_barrier)."""
        buf, hdl = self._p2p
        ny, Nk = self.N // self.nranks, self.N // 2 + 1
        st = self.eng.stream if stream is None else ct.c_void_p(stream.cuda_stream)
        _lib.call("apk_slab_transpose_p2p", self.eng._plan, _ptr(g), ct.c_void_p(hdl.buffer_ptrs_dev),
                  f * self.N * ny * Nk * 8, self.nranks, st)
        return buf[f]

    def p2p_barrier(self, channel: int) -> None:
        """Device-side barrier of all ranks on the current stream (symmetric-memory signal pads)."""
        self._p2p[1].barrier(channel=channel)

    # -- 7. binning ----------------------------------------------------------------------------------
    def make_binning(self, y0: int, ny: int, kmin, dk, kmax, compensation, interlaced):
        axes = {"ia": np.arange(self.N), "ib": np.arange(y0, y0 + ny), "key": ("slabT", y0, ny)}
        return self.eng.binning(kmin, dk, kmax, compensation, interlaced, axes=axes)

    def bin(self, binning, c1, c1s) -> torch.Tensor:
        return self.eng.bin_power_raw(binning, c1, c1s)

    def to_reduce_tensors(self, raw, total_mass) -> tuple:
        """([ksum | psum_re | psum_im | total mass] float64, mode counts int64): the two all-reduces of the path.
        The counts stay integers end to end."""
        nb1 = raw.shape[1]
        out = torch.empty(3 * nb1 + 1, dtype=torch.float64, device=self.device)
        out[: 3 * nb1] = raw[:3].reshape(-1)
        out[3 * nb1] = total_mass[0]
        return out, raw[3].view(torch.int64).clone()


def bind_host_to_gpu(device_index: int):
    """Restricts this process to the CPUs next to GPU ``device_index`` (NVML's affinity mask), so that the pinned host
    buffers it allocates afterwards -- first touch -- and its copy threads sit on the GPU's NUMA node.  One process per GPU
    uploads its share of the particles; without this, eight concurrent uploads cross the socket link.  Returns the number
    of CPUs kept, or None if NVML or the affinity call is unavailable (nothing changes then)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        words = (max(os.cpu_count() or 1, 1) + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * i + b for i, w in enumerate(mask) for b in range(64) if (int(w) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:  # noqa: BLE001 -- an optimisation only
        return None


class TorchDistComm:
    """The four exchanges of the path over torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""

    def __init__(self, group=None):
        self.group = group
        self.P = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    def _global(self, r: int) -> int:
        return r if self.group is None else dist.get_global_rank(self.group, r)

    def all_to_all_rows(self, send: torch.Tensor, send_counts: list, alloc) -> torch.Tensor:
        """all-to-all-v of the leading dimension; alloc(rows) makes the receive buffer."""
        sc = torch.tensor(send_counts, dtype=torch.int64, device=send.device)
        rc = torch.empty_like(sc)
        dist.all_to_all_single(rc, sc, group=self.group)
        recv_counts = rc.cpu().tolist()
        recv = alloc(int(sum(recv_counts)))
        dist.all_to_all_single(recv, send, output_split_sizes=recv_counts, input_split_sizes=send_counts,
                               group=self.group)
        return recv

    def ring_exchange(self, to_prev: torch.Tensor, to_next: torch.Tensor):
        """Send to_prev to rank-1 and to_next to rank+1; returns (from_next, from_prev)."""
        prev, nxt = self._global((self.rank - 1) % self.P), self._global((self.rank + 1) % self.P)
        from_next, from_prev = torch.empty_like(to_prev), torch.empty_like(to_next)
        # with two ranks prev == nxt: receives are posted in the order the peer sends (lo, then hi)
        ops = [dist.P2POp(dist.isend, to_prev, prev, self.group), dist.P2POp(dist.isend, to_next, nxt, self.group),
               dist.P2POp(dist.irecv, from_next, nxt, self.group), dist.P2POp(dist.irecv, from_prev, prev, self.group)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        return from_next, from_prev

    def all_to_all_blocks(self, send: torch.Tensor) -> torch.Tensor:
        """send[s] goes to rank s; returns recv with recv[q] = what rank q sent here (equal blocks)."""
        recv = torch.empty_like(send)
        if send.is_complex():
            dist.all_to_all_single(torch.view_as_real(recv), torch.view_as_real(send), group=self.group)
        else:
            dist.all_to_all_single(recv, send, group=self.group)
        return recv

    def all_reduce_sum(self, t: torch.Tensor) -> torch.Tensor:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t


class SlabPk:
    """P(k) of particles distributed over the ranks of ``comm`` (default: the torch.distributed world)."""

    def __init__(self, Nmesh: int, BoxSize: float, resampler: str = "tsc", interlaced: bool = False,
                 compensated: bool = False, device=None, group=None, backend=None, comm=None):
        if comm is None and group is None and dist.is_initialized() and dist.get_world_size() > 1 \
                and dist.get_backend() == "nccl" and os.environ.get("APK_SLAB_HP_NCCL", "1") != "0":
            # NCCL's kernels run on its own internal stream.  At default priority they would queue behind the
            # deposit's grid, and so would everything the side stream does after them: a communicator with
            # high-priority streams lets the exchanges run WHILE the deposit kernels do.
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
            group = dist.new_group(ranks=list(range(dist.get_world_size())), backend="nccl", pg_options=opts)
        self.comm = comm if comm is not None else TorchDistComm(group)
        self.P, self.rank = self.comm.P, self.comm.rank
        self.N, self.L = int(Nmesh), float(BoxSize)
        if self.N % self.P:
            raise AstrildPkError(f"Nmesh {self.N} must be divisible by the number of ranks {self.P}")
        self.n0 = self.N // self.P
        self.x0 = self.rank * self.n0
        self.ny, self.y0 = self.n0, self.x0              # after the transpose this rank owns these y
        self.Nk = self.N // 2 + 1
        self.resampler, self.interlaced, self.compensated = str(resampler).lower(), bool(interlaced), bool(compensated)
        if backend is None:
            backend = CudaSlabBackend(self.N, self.L, self.x0, self.n0, self.P, device)
        self.backend = backend
        self.eng = getattr(backend, "eng", None)
        self.ghost_lo, self.ghost_hi = 1, 2
        self.profile, self.last_profile = False, {}
        self.last_info: dict = {}                                 # which transpose ran, its time, the critical stage
        self.stream_rows = 1 << 24                                # host inputs larger than this are uploaded in chunks
        self.p2p = os.environ.get("APK_SLAB_P2P", "1") != "0"     # fused peer-store transpose (falls back to NCCL)
        if self.P > 1 and self.n0 < 2:
            raise AstrildPkError("each rank needs at least 2 mesh planes")

    # ------------------------------------------------------------------ helpers
    def lattice_planes(self, n: int) -> tuple:
        """Planes [a, b) of an n^3 lattice this rank generates (contiguous, near its own slab)."""
        a = (self.rank * n) // self.P
        b = ((self.rank + 1) * n) // self.P
        return a, b

    def _exchange_ghosts(self, meshes: list) -> list:
        """Send the ghost planes of every mesh to the ring neighbours (one batch of sends and receives for all
        of them), add what arrives; returns the owned planes."""
        lo, hi, n0 = self.ghost_lo, self.ghost_hi, self.n0
        if self.P == 1:
            return list(meshes)                          # whole periodic mesh: the kernel wrapped already
        # ghost_lo is the previous rank's last plane, ghost_hi the next rank's first planes
        to_prev = torch.stack([m[:lo] for m in meshes]) if len(meshes) > 1 else meshes[0][:lo].contiguous()
        to_next = torch.stack([m[lo + n0:] for m in meshes]) if len(meshes) > 1 else meshes[0][lo + n0:].contiguous()
        from_next, from_prev = self.comm.ring_exchange(to_prev, to_next)
        if len(meshes) == 1:
            from_next, from_prev = from_next[None], from_prev[None]
        owned = []
        for f, m in enumerate(meshes):
            o = m[lo: lo + n0]
            self.backend.accumulate(o[n0 - lo:], from_next[f])
            self.backend.accumulate(o[:hi], from_prev[f])
            owned.append(o)
        return owned

    def _transpose(self, grids: list) -> list:
        """[n0][N][Nk] x-slabs -> [N][ny][Nk] y-slabs; one all-to-all per field so that the
        received blocks (ordered by source rank = by x) already are the transposed slab."""
        P, n0, ny, Nk = self.P, self.n0, self.ny, self.Nk
        out = []
        for g in grids:
            # block for rank s: [x_local][y in s's range][z]
            send = g.reshape(n0, P, ny, Nk).permute(1, 0, 2, 3).contiguous()     # [P][n0][ny][Nk]
            recv = send if P == 1 else self.comm.all_to_all_blocks(send)
            out.append(recv.reshape(self.N, ny, Nk))
        return out

    # ------------------------------------------------------------------ the path
    def power(self, pos, mass=None, pos_scale: float | None = None, kmin: float = 0.0, dk=None, kmax=None,
              normalize: bool = True, routed: bool = False, mode: str = "1d", Nmu: int | None = None, poles=(),
              los=(0.0, 0.0, 1.0)) -> dict:
        """(k, power, modes) of this rank's particles together with everyone else's.

        pos: this rank's share, (Np,3) or three columns, host or device.  routed=True promises that
        every particle already sits on the rank owning floor(g_x) (skips the all-to-all-v).
        mode='2d' (or poles): (k, mu) wedges and multipoles as FFTPower(mode='2d', Nmu=, poles=, los=) returns them --
        ``k``, ``mu``, ``power``, ``modes`` of shape (Nk, Nmu) and ``poles`` (CUDA backend only).
        """
        if mode not in ("1d", "2d"):
            raise AstrildPkError("mode must be '1d' or '2d'")
        self._kmu = None
        if mode == "2d" or len(poles):
            self._kmu = (1 if mode == "1d" else (5 if Nmu is None else int(Nmu)), tuple(int(e) for e in poles),
                         tuple(float(x) for x in los), mode)
        be, P = self.backend, self.P
        ps = 1.0 / self.L if pos_scale is None else float(pos_scale)
        marks = []

        def mark(name):
            if self.profile and torch.cuda.is_available():
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                marks.append((name, ev))

        mark("start")
        scalar_mass = 1.0
        if mass is not None and np.isscalar(mass):   # one weight for all: deposit unit weights, scale at the end
            scalar_mass, mass = float(mass), None
        eng = getattr(be, "eng", None)
        self.last_info = {"transpose": "local" if P == 1 else "nccl-a2a"}
        self._transpose_events = []
        if eng is not None:
            first = pos[0] if isinstance(pos, (tuple, list)) else pos
            on_host = not (isinstance(first, torch.Tensor) and first.is_cuda)
            if on_host and P > 1 and not routed and int(first.shape[0]) > self.stream_rows and hasattr(be, "setup_p2p"):
                # host inputs: chunked upload, each chunk routed and deposited while the next one is in flight
                grids, total = self._streamed_host(pos, mass, ps, mark)
                return self._finish(grids, total, kmin, dk, kmax, normalize, marks, mark, scalar_mass)
            # ONE upload serves routing and deposit
            pos = tuple(eng._to_device(c) for c in pos) if isinstance(pos, (tuple, list)) else eng._to_device(pos)
            if mass is not None:
                mass = eng._to_device(mass)
        side = getattr(be, "side_stream", None)
        if (P > 1 and self.interlaced and self.p2p and side is not None and torch.cuda.is_available()
                and isinstance(self.comm, TorchDistComm) and hasattr(be, "setup_p2p")
                and be.setup_p2p(self.comm.group, 2)):
            grids, total = self._pipelined_pair(pos, mass, ps, routed, mark)
            return self._finish(grids, total, kmin, dk, kmax, normalize, marks, mark, scalar_mass)
        # 1. route: only particles that change slab travel; the slab deposit ignores particles it
        #    does not own, so the caller's arrays are deposited as they are
        side = getattr(be, "side_stream", None)
        overlap = side is not None and torch.cuda.is_available()
        incoming = None                              # (positions, masses, event) of the particles other ranks send
        if not (routed or P == 1):
            sp, sm, counts = be.route(pos, mass, ps)
            mark("route")
            if overlap:
                # the all-to-all-v runs on the side stream while the own particles are deposited (issued below,
                # BEFORE the host blocks on the exchanged counts)
                main = torch.cuda.current_stream(be.device)
                routed_ev = torch.cuda.Event()
                routed_ev.record(main)
        # 2./3. deposit + ghosts
        pair_mode = self.interlaced and hasattr(be, "deposit_pair")
        shifts = (0.0, 0.5) if self.interlaced else (0.0,)

        def deposit_into(rp, rm, meshes, method="auto"):
            if pair_mode:
                return list(be.deposit_pair(rp, rm, self.resampler, ps, out=None if meshes is None else tuple(meshes),
                                            **({} if method == "auto" else {"method": method})))
            return [be.deposit(rp, rm, self.resampler, sh, ps, out=None if meshes is None else meshes[i],
                               **({} if method == "auto" else {"method": method})) for i, sh in enumerate(shifts)]

        if routed or P == 1:
            meshes = deposit_into(pos, mass, None)
        elif overlap and hasattr(be, "zeroed_meshes"):
            # own particles on the main stream; meanwhile, on the side stream, the leavers are exchanged and --
            # a small set near the slab faces -- added with plain REDs into the same meshes
            meshes = be.zeroed_meshes(len(shifts))
            zeroed = torch.cuda.Event()
            zeroed.record(main)
            deposit_into(pos, mass, meshes)
            with torch.cuda.stream(side):
                side.wait_event(routed_ev)
                fp = self.comm.all_to_all_rows(sp, counts, lambda rows: be.empty_like_rows(sp, rows))
                fm = (self.comm.all_to_all_rows(sm, counts, lambda rows: be.empty_like_rows(sm, rows))
                      if sm is not None else None)
                small = fp.shape[0] < (1 << 22)
                if small and fp.shape[0]:
                    side.wait_event(zeroed)
                    deposit_into(fp, fm, meshes, method="atomic")
                arrived = torch.cuda.Event()
                arrived.record(side)
            main.wait_event(arrived)
            for t in (fp, fm, sp, sm):
                if t is not None:
                    t.record_stream(main)
                    t.record_stream(side)
            if not small:
                deposit_into(fp, fm, meshes)
        else:
            meshes = deposit_into(pos, mass, None)
            fp = self.comm.all_to_all_rows(sp, counts, lambda rows: be.empty_like_rows(sp, rows))
            fm = (self.comm.all_to_all_rows(sm, counts, lambda rows: be.empty_like_rows(sm, rows))
                  if sm is not None else None)
            if fp.shape[0]:
                deposit_into(fp, fm, meshes)
        mark("deposit")
        grids, total = self._ghosts_fft_transpose(meshes, mark)
        del meshes
        return self._finish(grids, total, kmin, dk, kmax, normalize, marks, mark, scalar_mass)

    def _finish_kmu(self, grids, total, kmin, dk, kmax, normalize, comp, scalar_mass: float) -> dict:
        """(k, mu) wedges and multipoles on the transposed slab (row N4): every rank bins its y-range, the float64 sums
        and the integer mode counts are all-reduced separately, the tail is nbodykit's project_to_basis."""
        Nmu, poles, los, mode = self._kmu
        eng = getattr(self.backend, "eng", None)
        if eng is None:
            raise AstrildPkError("the (k, mu) mode needs the CUDA backend")
        axes = {"ia": np.arange(self.N), "ib": np.arange(self.y0, self.y0 + self.ny), "key": ("slabT", self.y0, self.ny)}
        kb = eng.kmu_binning(kmin, dk, kmax, Nmu, poles, los, comp, self.interlaced, axes=axes)
        raw = eng.bin_kmu_raw(kb, grids[0], grids[1] if self.interlaced else None)
        red = torch.cat([raw.reshape(-1), total.reshape(-1)[:1].to(torch.float64)])
        cnt = raw[2].view(torch.int64).clone()
        if self.P > 1:
            red = self.comm.all_reduce_sum(red)
            cnt = self.comm.all_reduce_sum(cnt)
        host = red.cpu().numpy()
        W = host[-1] * scalar_mass
        N, L = self.N, self.L
        field_scale = (N ** 3 / W) * scalar_mass if normalize else scalar_mass / (L / N) ** 3
        res = eng.finish_kmu(host[:-1].reshape(raw.shape), cnt.cpu().numpy().astype(np.int64), kb,
                             L ** 3 * field_scale ** 2 / float(N) ** 6)
        res["total_mass"] = W
        if mode == "1d":                                  # FFTPower(mode='1d', poles=...): the 1-D spectrum plus the poles
            res = {"k": res["poles"]["k"], "power": res["poles"]["power_0"], "modes": res["poles"]["modes"],
                   "edges": res["edges"], "poles": res["poles"], "total_mass": W}
        return res

    def _timed_transpose(self, fn, stream):
        """Runs fn() between two CUDA events on `stream` (the transposes' own time, for the NVLink roofline)."""
        if not (self.profile and torch.cuda.is_available()):
            return fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        out = fn()
        e1.record(stream)
        self._transpose_events.append((e0, e1))
        return out

    def _ghosts_fft_transpose(self, meshes: list, mark) -> tuple:
        """Stages 3-6 for complete slab meshes: ghost planes, 2-D r2c, x<->y transpose, 1-D c2c.
        Returns (transposed k-grids [N][ny][Nk], total mass tensor)."""
        be, P = self.backend, self.P
        side = getattr(be, "side_stream", None)
        owned = self._exchange_ghosts(meshes)
        mark("ghosts")
        # 4. 2-D FFT, 5. transpose, 6. 1-D FFT
        #    pipelined per field: while field f is packed and exchanged on a side stream, the 2-D FFT of
        #    field f+1 (and later the 1-D FFT of field f-1) runs on the main stream
        use_p2p = (P > 1 and self.p2p and isinstance(self.comm, TorchDistComm) and hasattr(be, "setup_p2p")
                   and be.setup_p2p(self.comm.group, 2 if self.interlaced else 1))
        if use_p2p:
            self.last_info["transpose"] = "p2p-store"
        elif P > 1 and self.p2p and hasattr(be, "_p2p_error"):
            self.last_info["transpose_fallback_reason"] = be._p2p_error
        has_dc = hasattr(be, "dc_sum")
        total = None
        if use_p2p and side is not None:
            # peer-store transposes on the side stream: field f travels over NVLink while the main stream
            # transforms field f+1 (2-D) and, later, field f-1 (1-D)
            main = torch.cuda.current_stream(be.device)
            landed = []
            for f, o in enumerate(owned):
                g2 = be.fft2d(o)
                if f == 0:
                    total = be.dc_sum(g2) if has_dc else be.mesh_sum(o)
                ready = torch.cuda.Event()
                ready.record(main)
                o.record_stream(side)                # the side stream reads these planes
                with torch.cuda.stream(side):
                    side.wait_event(ready)
                    if f == 0:
                        be.p2p_barrier(0)            # every rank is done reading the previous step's buffers
                    self._timed_transpose(lambda: (be.transpose_p2p_store(f, g2, side), be.p2p_barrier(1 + (f & 1))), side)   # field f has landed everywhere
                    ev = torch.cuda.Event()
                    ev.record(side)
                landed.append(ev)
            grids = []
            for f, ev in enumerate(landed):
                main.wait_event(ev)
                grids.append(be.fft1d(be._p2p[0][f], self.ny))
            del owned
            mark("fft+transpose")
        elif use_p2p or side is None or P == 1 or len(owned) == 1:
            grids = [be.fft2d(o) for o in owned]
            total = be.dc_sum(grids[0]) if has_dc else be.mesh_sum(owned[0])
            mark("fft2d")
            cur = torch.cuda.current_stream(be.device) if torch.cuda.is_available() and hasattr(be, "device") else None
            grids = self._timed_transpose(lambda: be.transpose_p2p(grids) if use_p2p else self._transpose(grids), cur) \
                if cur is not None else (be.transpose_p2p(grids) if use_p2p else self._transpose(grids))
            del owned
            mark("transpose")
            grids = [be.fft1d(g, self.ny) for g in grids]
            mark("fft1d")
        else:
            main = torch.cuda.current_stream(be.device)
            moved, done = [], []
            for f, o in enumerate(owned):
                g2 = be.fft2d(o)
                if f == 0:
                    total = be.dc_sum(g2) if has_dc else be.mesh_sum(o)
                ready = torch.cuda.Event()
                ready.record(main)
                o.record_stream(side)
                with torch.cuda.stream(side):
                    side.wait_event(ready)
                    t = self._timed_transpose(lambda: self._transpose([g2])[0], side)
                    ev = torch.cuda.Event()
                    ev.record(side)
                t.record_stream(main)
                moved.append(t)
                done.append(ev)
            grids = []
            for t, ev in zip(moved, done):
                main.wait_event(ev)
                grids.append(be.fft1d(t, self.ny))
            del owned
            mark("fft+transpose")
        return grids, total

    def _streamed_host(self, pos, mass, ps: float, mark) -> tuple:
        """HOST inputs on P > 1 GPUs: the rank's share is uploaded in chunks on a copy stream; every chunk is routed
        (its leavers staged on the device) and deposited into the slab meshes while the next chunk is in flight, so the
        end-to-end time is the PCIe time plus the tail (exchange of the leavers, ghosts, FFTs, binning) instead of
        upload + everything.  Pinned host memory gives asynchronous copies.  Returns (k-grids, total mass tensor)."""
        be, eng = self.backend, self.backend.eng
        cols = list(pos) if isinstance(pos, (tuple, list)) else None

        def host_tensor(a):
            t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
            return t if t.dtype in (torch.float32, torch.float64) else t.to(torch.float64)

        srcs = [host_tensor(c) for c in cols] if cols is not None else [host_tensor(pos)]
        if mass is not None:
            srcs.append(host_tensor(mass).to(srcs[0].dtype))
        npart, rows = int(srcs[0].shape[0]), self.stream_rows
        nchunk = (npart + rows - 1) // rows
        main = torch.cuda.current_stream(be.device)
        copy_stream = torch.cuda.Stream(be.device)
        be.prepare_ffts(self.ny)
        nshift = 2 if self.interlaced else 1
        meshes = be.zeroed_meshes(nshift)
        bufs = [[torch.empty((rows,) + tuple(t.shape[1:]), dtype=t.dtype, device=be.device) for t in srcs] for _ in range(2)]
        copy_stream.wait_stream(main)
        for pair in bufs:
            for buf in pair:
                buf.record_stream(copy_stream)
        counts = torch.zeros((nchunk, 2 * self.P), dtype=torch.int64, device=be.device)
        ready, free, staged = [torch.cuda.Event() for _ in range(2)], [None, None], []
        for c, a in enumerate(range(0, npart, rows)):
            b, sl = min(a + rows, npart), c & 1
            with torch.cuda.stream(copy_stream):
                if free[sl] is not None:
                    copy_stream.wait_event(free[sl])
                for buf, src in zip(bufs[sl], srcs):
                    buf[: b - a].copy_(src[a:b], non_blocking=True)
                ready[sl].record(copy_stream)
            main.wait_event(ready[sl])
            part = [buf[: b - a] for buf in bufs[sl]]
            ppos = tuple(part[:3]) if cols is not None else part[0]
            pm = part[-1] if mass is not None else None
            # every particle of the chunk may leave: a staging buffer of the chunk's size never needs a second pass
            staged.append(be.route_begin(ppos, pm, ps, capacity=b - a, counts=counts[c]))
            if self.interlaced:
                be.deposit_pair(ppos, pm, self.resampler, ps, out=tuple(meshes))
            else:
                be.deposit(ppos, pm, self.resampler, 0.0, ps, out=meshes[0])
            free[sl] = torch.cuda.Event()
            free[sl].record(main)
        main.wait_stream(copy_stream)
        mark("upload+deposit")
        # the leavers of all chunks, grouped by destination, in ONE all-to-all-v
        per = counts[:, : self.P].cpu().numpy()                      # [chunk][destination]
        send_counts = per.sum(axis=0).astype(np.int64).tolist()
        segs_p, segs_m = [], []
        for d in range(self.P):
            for c, h in enumerate(staged):
                n = int(per[c, d])
                if n:
                    o = int(per[c, :d].sum())
                    segs_p.append(h["out_pos"][o:o + n])
                    if h["out_mass"] is not None:
                        segs_m.append(h["out_mass"][o:o + n])
        like = staged[0]["out_pos"]
        sp = torch.cat(segs_p) if segs_p else like[:0]
        sm = (torch.cat(segs_m) if segs_m else staged[0]["out_mass"][:0]) if mass is not None else None
        del staged
        fp = self.comm.all_to_all_rows(sp, send_counts, lambda r: be.empty_like_rows(sp, r))
        fm = self.comm.all_to_all_rows(sm, send_counts, lambda r: be.empty_like_rows(sm, r)) if sm is not None else None
        if fp.shape[0]:
            method = "atomic" if fp.shape[0] < (1 << 22) else "auto"
            if self.interlaced:
                be.deposit_pair(fp, fm, self.resampler, ps, out=tuple(meshes), method=method, timed=False)
            else:
                be.deposit(fp, fm, self.resampler, 0.0, ps, out=meshes[0], method=method)
        mark("exchange")
        return self._ghosts_fft_transpose(meshes, mark)

    def routing_stress(self, pos, pos_scale: float, kmin: float, seed: int = 0, steps: int = 2) -> dict:
        """The all-to-all-v at its design load: the rank's particles are shifted along x by (i mod P) / P, i = index in
        blocks of 65536, so that (P - 1) / P of them change slab (each rank then holds a 1/P sample of every slab) while
        the order inside a block stays coherent.  Runs the whole path on that set and returns the step time and the
        stage split of rank 0's last step; not part of the timed region of bench.py."""
        P = self.P
        cols = [c.clone() for c in pos]
        idx = torch.arange(cols[0].numel(), device=cols[0].device) // 65536
        shift = (idx % P).to(cols[0].dtype) / P
        x = cols[0] * (pos_scale if pos_scale != 1.0 else 1.0) + shift
        cols[0] = ((x - torch.floor(x)) / (pos_scale if pos_scale != 1.0 else 1.0)).clamp_(0, torch.finfo(cols[0].dtype).max)
        cols[0] = torch.where(cols[0] * pos_scale >= 1.0, torch.zeros_like(cols[0]), cols[0])
        del idx, shift, x
        was = self.profile
        self.profile = True
        res = None
        times = []
        for _ in range(steps):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            res = self.power(tuple(cols), pos_scale=pos_scale, kmin=kmin, normalize=True)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        self.profile = was
        t = torch.tensor([times[-1]], dtype=torch.float64, device=cols[0].device)
        if P > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.comm.group)
        moved = cols[0].numel() * (P - 1) / P
        return {"ms_per_step": float(t.item()), "particles_leaving_per_rank": int(moved),
                "bytes_out_per_gpu": int(moved * 12), "stages_ms_rank0": {k: round(v, 3) for k, v in self.last_profile.items()},
                "modes0": int(res["modes"][0])}

    def _pipelined_pair(self, pos, mass, ps: float, routed: bool, mark):
        """Interlaced twins on P > 1 GPUs with peer-mapped transposes: everything that concerns mesh 0 after its
        tile kernel -- ghost planes, 2-D FFT, peer-store transpose -- runs on the side stream WHILE the main
        stream still deposits mesh 1 (an issue-bound kernel; the side work is bandwidth- and NVLink-bound), and
        so do the particle exchange and the (few) received particles, which are added with plain REDs.
        Returns (transposed k-grids [N][ny][Nk] of both meshes, total mass tensor)."""
        be = self.backend
        main, side = torch.cuda.current_stream(be.device), be.side_stream
        self.last_info["transpose"] = "p2p-store"
        be.prepare_ffts(self.ny)                     # plans first: they size the workspace once
        sp = sm = fp = fm = None
        if not routed:
            pending = be.route_begin(pos, mass, ps)  # the counts are read later, on the side stream: no idle GPU
        meshes = be.zeroed_meshes(2)
        for m in meshes:
            m.record_stream(side)
        begun = torch.cuda.Event()
        begun.record(main)                           # meshes zeroed, leavers staged
        first = torch.cuda.Event()
        first.record(main)                           # (creates the handle; re-recorded inside the library)
        be.set_first_mesh_event(first)
        try:
            be.deposit_pair(pos, mass, self.resampler, ps, out=tuple(meshes))
        finally:
            be.set_first_mesh_event(None)
        both = torch.cuda.Event()
        both.record(main)
        mesh0_ready = first
        if not routed:
            with torch.cuda.stream(side):
                side.wait_event(begun)
                sp, sm, counts = be.route_end(pending)
                fp = self.comm.all_to_all_rows(sp, counts, lambda rows: be.empty_like_rows(sp, rows))
                fm = (self.comm.all_to_all_rows(sm, counts, lambda rows: be.empty_like_rows(sm, rows))
                      if sm is not None else None)
                small = fp.shape[0] < (1 << 22)
                if small and fp.shape[0]:
                    be.deposit_pair(fp, fm, self.resampler, ps, out=tuple(meshes), method="atomic", timed=False)
                arrived = torch.cuda.Event()
                arrived.record(side)
            for t in (fp, fm, sp, sm):
                if t is not None:
                    t.record_stream(main)
                    t.record_stream(side)
            if not small:                            # many leavers: sorted deposit after the own particles
                main.wait_event(arrived)
                be.deposit_pair(fp, fm, self.resampler, ps, out=tuple(meshes))
                both = torch.cuda.Event()
                both.record(main)
                mesh0_ready = both
        mark("deposit")
        landed = []
        with torch.cuda.stream(side):
            for f in range(2):
                side.wait_event(mesh0_ready if f == 0 else both)
                owned = self._exchange_ghosts([meshes[f]])[0]
                g2 = be.fft2d(owned)
                if f == 0:
                    total = be.dc_sum(g2)
                    total.record_stream(main)
                    be.p2p_barrier(0)                # every rank is done reading the previous step's buffers
                self._timed_transpose(lambda: (be.transpose_p2p_store(f, g2, side), be.p2p_barrier(1 + f)), side)   # field f has landed everywhere
                ev = torch.cuda.Event()
                ev.record(side)
                landed.append(ev)
        grids = []
        for f, ev in enumerate(landed):
            main.wait_event(ev)
            # (field 0's 1-D transform on the side stream, under the twin's deposit, was measured: the tail shrinks by
            # 0.1 ms and the deposit grows by 0.3 ms at 8 GPUs -- it stays here)
            grids.append(be.fft1d(be._p2p[0][f], self.ny))
        del meshes
        mark("ghosts+fft+transpose")
        return grids, total

    def _finish(self, grids, total, kmin, dk, kmax, normalize, marks, mark, scalar_mass: float = 1.0) -> dict:
        be, P = self.backend, self.P
        # 7. binning on the transposed slab
        comp = (self.resampler, self.interlaced) if self.compensated else None
        if getattr(self, "_kmu", None) is not None:
            return self._finish_kmu(grids, total, kmin, dk, kmax, normalize, comp, scalar_mass)
        binning = be.make_binning(self.y0, self.ny, kmin, dk, kmax, comp, self.interlaced)
        raw = be.bin(binning, grids[0], grids[1] if self.interlaced else None)
        mark("bin")
        # 8. reduce: float64 sums in one all-reduce, the integer mode counts in another
        red, cnt = be.to_reduce_tensors(raw, total)
        if P > 1:
            red = self.comm.all_reduce_sum(red)
            cnt = self.comm.all_reduce_sum(cnt)
        host = red.cpu().numpy()
        nsum = cnt.cpu().numpy().astype(np.int64)
        mark("reduce")
        if marks:
            prof: dict = {}
            for (_, e0), (name, e1) in zip(marks[:-1], marks[1:]):
                prof[name] = prof.get(name, 0.0) + e0.elapsed_time(e1)
            self.last_profile = prof
            if prof:
                self.last_info["critical_path"] = max(prof, key=prof.get)
        tev = getattr(self, "_transpose_events", None)
        if tev:
            self.last_info["transpose_ms"] = float(sum(a.elapsed_time(b) for a, b in tev))
            self._transpose_events = None
        nb1 = (len(host) - 1) // 3
        ksum, pre, pim = host[:nb1], host[nb1:2 * nb1], host[2 * nb1:3 * nb1]
        W = host[3 * nb1] * scalar_mass
        N, L = self.N, self.L
        # the meshes hold unit-weight deposits when the caller's weight is one scalar: it enters here
        field_scale = (N ** 3 / W) * scalar_mass if normalize else scalar_mass / (L / N) ** 3
        scale = L ** 3 * field_scale ** 2 / float(N) ** 6
        with np.errstate(invalid="ignore", divide="ignore"):
            k = (ksum / nsum)[1:-1]
            power = ((pre + 1j * pim) * scale / nsum)[1:-1]
        return {"k": k, "power": power, "modes": nsum[1:-1].copy(), "edges": binning.edges, "Nsum": nsum,
                "total_mass": W}
