"""Operator algebra and the ``integrate`` driver — the reference-facing plugin surface.

Same names, argument meaning and error behaviour as the reference's operator layer for the hot
path (``torchfsm/operator/_base.py``): operators are sums of generator terms, ``integrate(u_0,
u_0_fft, dt, step, mesh, progressive, trajectory_recorder, return_in_fourier)`` advances a
state, ``__call__`` evaluates the right-hand side, ``set_integrator`` picks the scheme.
What differs is what ``_build_integrator`` installs: instead of a torch integrator object the
operator is lowered to a *fused step program* executed by the CUDA library through the C ABI
(``include/fsm_b200.h``). Anything the fused programs cannot express raises
``NotImplementedError`` — there is no torch/CPU fallback on this path.
"""
import ctypes
import os
from typing import Callable, List, Optional, Sequence

import torch

from . import _cabi
from .autograd import GradientMode, _adjoint_terms
from .integrator import (ETDRKIntegrator, SETDRKIntegrator, RKIntegrator, RK_TABLEAUS, integrator_name, build_tables,
                         has_imag)
from .unrolled import PairedSpectrumStepper
from .mesh import FourierMesh, MeshGrid

# "linear_tensor": a ready-made L(k) tensor in the reference layout (what a genuine torchfsm operator registered,
# reference_adapter.lower); time stepping only
_LINEAR_KINDS = ("laplacian", "biharmonic", "spatial_derivative", "implicit_unit_source", "linear_tensor")
# channel-changing cores that are pure symbol products: evaluated by the point-wise spectral map (fsm_spectral_map)
_MAP_KINDS = ("grad", "div", "curl", "vorticity2velocity")
# diagnostics composed of a map, one convection evaluation and a pressure solve
_COMPOSITE_KINDS = ("velocity2pressure", "vorticity2pressure")
# nonlinear cores without a fused program: evaluated by composing the library's passes on the host (c2r, a point-wise
# physical-space function, r2c, spectral map) and handed to the integrator stage (fsm_stage_combine)
# "custom_nonlinear": a user-defined NonlinearFunc (the reference's core protocol, operator/_base.py:56-103) called on
# full spectra that the library's passes produce and consume
_EXTERNAL_KINDS = ("implicit_func_source", "conservative_convection", "custom_nonlinear")
_PROGRAM_OF = {"convection": _cabi.PROG_CONVECTION, "ks_convection": _cabi.PROG_KS,
               "vorticity_convection": _cabi.PROG_NS2D_VORT, "ns_pressure_convection": _cabi.PROG_NS3D}


class _Term:
    __slots__ = ("kind", "coef", "params")

    def __init__(self, kind, coef=1, params=None):
        self.kind, self.coef, self.params = kind, coef, dict(params or {})

    def scaled(self, s):
        if isinstance(s, torch.Tensor) and not isinstance(self.coef, torch.Tensor) and self.coef == 1:
            return _Term(self.kind, s, self.params)        # the parameter itself: no graph node made at construction
        return _Term(self.kind, self.coef * s, self.params)


# ------------------------------------------------------------------------------------------------
# rot-half layout helpers (see include/fsm_b200.h)
# ------------------------------------------------------------------------------------------------
def _per_sample(coef: torch.Tensor) -> bool:
    """A tensor-valued coefficient the nonlinear terms accept: one value, or one value per sample ((B, 1, 1, ..))."""
    return coef.dim() == 0 or coef.numel() == coef.shape[0]


def _rot_half(t: torch.Tensor, shape) -> torch.Tensor:
    """(X, n0[, n1[, n2]]) full-layout tensor -> (X, rot-half modes) contiguous."""
    nd = len(shape)
    nh = shape[-1] // 2 + 1
    t = t[..., :nh]
    if nd == 2:
        t = t.permute(0, 2, 1)
    elif nd == 3:
        t = t.permute(0, 2, 3, 1)
    return t.contiguous()


def _expand_table(t: torch.Tensor, shape) -> torch.Tensor:
    """Broadcastable (1|B, 1|C, ...) reference-layout table -> (Bt, Ct, *shape) with Bt in {1, B}, Ct in {1, C}."""
    return t.expand(t.shape[0], t.shape[1], *shape)


def _same_device(a: torch.device, b: torch.device) -> bool:
    if a.type != b.type:
        return False
    if a.type != "cuda":
        return True
    ia = a.index if a.index is not None else torch.cuda.current_device()
    ib = b.index if b.index is not None else torch.cuda.current_device()
    return ia == ib


class _StreamWork:
    """Completion handle of work queued on the (side) stream that was current at construction; ``wait()`` orders
    the then-current stream after it, like the Work objects of torch.distributed."""

    def __init__(self, device):
        self._ev = None
        if device.type == "cuda":
            self._ev = torch.cuda.Event()
            self._ev.record(torch.cuda.current_stream(device))

    def wait(self):
        if self._ev is not None:
            torch.cuda.current_stream().wait_event(self._ev)


class _GuardedLib:
    """The C library acts on the CURRENT CUDA device (twiddle tables are filled per device, kernels launch on
    it), while a plan belongs to the device of its mesh: every entry point is therefore called with that device
    made current, like torch ops follow their tensors' device."""

    def __init__(self, lib, device):
        self._lib, self._device = lib, device

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        if self._device.type != "cuda":
            return fn
        dev = self._device

        def call(*args):
            with torch.cuda.device(dev):
                return fn(*args)
        return call


class FusedStepper:
    """What ``_build_integrator`` installs: a plan of the CUDA library plus the buffers it needs.

    Exposes the reference's integrator protocol on full-spectrum tensors (``.dt``, ``.step(u_hat)``,
    ``.forward(u_hat, dt)``; operator/_base.py:462-491) and the native half-spectrum entry points
    used by ``integrate``.
    """

    def __init__(self, f_mesh: FourierMesh, batch: int, n_channel: int, program: int, integrator: str, dt: float,
                 linear_coef: Optional[torch.Tensor], nl_coef: float, source_hat: Optional[torch.Tensor],
                 kmax: Sequence[int], ks_remove_mean: bool, integrator_cfg: dict, chunk: int = 0,
                 tables: Optional[dict] = None, slab=None, lanes: int = 0, allocate: bool = True,
                 force_hat: Optional[torch.Tensor] = None, dynamic_force: bool = False,
                 nl_coef_b: Optional[torch.Tensor] = None):
        lib = _GuardedLib(_cabi.lib(), f_mesh.device)
        # slab = (rank, nranks, process_group): ONE 3-D grid decomposed over nranks GPUs
        self.slab = slab
        self.P, self.rank, self.group = (slab[1], slab[0], slab[2]) if slab else (1, 0, None)
        self.f_mesh, self.B, self.C, self.dt = f_mesh, batch, n_channel, dt
        self.shape = tuple(f_mesh.shape)
        self.n_dim = len(self.shape)
        self.device, self.rdtype, self.cdtype = f_mesh.device, f_mesh.dtype, f_mesh.cdtype
        if self.device.type != "cuda" and not _cabi.is_emulator():
            raise RuntimeError("torchfsm_b200 runs on CUDA devices only (tensor is on %s)" % self.device)
        nh = self.shape[-1] // 2 + 1
        self.nmodes = nh
        for n in self.shape[:-1]:
            self.nmodes *= n
        if self.P > 1:
            if self.n_dim != 3 or self.shape[0] % self.P or self.shape[1] % self.P:
                raise ValueError("slab decomposition needs a 3-D grid whose first two axes are divisible by the rank count")
            self.kyl, self.nxl = self.shape[1] // self.P, self.shape[0] // self.P
            self.nmodes //= self.P                        # local ky slab [n1/P][nh][n0]
            self.local_shape = (self.nxl,) + self.shape[1:]   # local physical x slab
        else:
            self.local_shape = self.shape
        self.integrator = integrator
        self._keep = []  # tensors referenced by the plan descriptor
        if self.rdtype == torch.float64 and not _cabi.is_emulator():
            # line buffers of one CTA must fit the 227 KB of shared memory of an SM (8 lines x 2 buffers, 16 B/point)
            if max(self.shape) > 512 and self.n_dim > 1:      # 1-D kernels hold one line per CTA
                raise NotImplementedError("fp64 grids are limited to 512 points per axis on the fused CUDA path "
                                          "(shared memory per SM); use fp32 or a smaller grid")
            if self.n_dim == 3 and program in (_cabi.PROG_CONVECTION, _cabi.PROG_NS3D) and self.shape[-1] > 256:
                raise NotImplementedError("fp64 3-D convection is limited to 256 points along the last axis on the "
                                          "fused CUDA path (shared memory per SM); use fp32 or a smaller grid")

        def real_table(t):
            t = _expand_table(t, self.shape)
            if t.shape[0] not in (1, batch):
                raise ValueError("a batched coefficient must have one entry per sample")
            if self.complex_tables:
                t = t.to(self.cdtype)                 # every table carries complex entries (include/fsm_b200.h)
            elif t.is_complex():
                if has_imag(t):
                    raise NotImplementedError(
                        "complex linear coefficients (odd-order linear terms) are supported on 1-D grids only")
                t = t.real
            if t.shape[1] > 1 and bool((t == t[:, :1]).all()):
                t = t[:, :1]
            return t if self.complex_tables else t.to(self.rdtype)

        desc = _cabi.FsmDesc()
        desc.struct_size = ctypes.sizeof(_cabi.FsmDesc)
        desc.dtype = _cabi.FSM_F32 if self.rdtype == torch.float32 else _cabi.FSM_F64
        desc.ndim = self.n_dim
        for i in range(3):
            desc.n[i] = self.shape[i] if i < self.n_dim else 1
            desc.kmax[i] = int(kmax[i]) if i < self.n_dim else 0
        desc.batch, desc.channels = batch, n_channel
        desc.program = program
        desc.integrator = _cabi.INTEGRATOR_IDS[integrator]
        desc.ks_remove_mean = 1 if ks_remove_mean else 0
        desc.chunk = int(chunk)
        desc.lanes = int(lanes)
        desc.dt = float(dt)
        desc.nl_coef = float(nl_coef)
        desc.dynamic_force = 1 if dynamic_force else 0
        if nl_coef_b is not None:             # tensor-valued coefficient of the convective term: one value per sample
            nl_coef_b = nl_coef_b.to(device=f_mesh.device, dtype=self.rdtype).reshape(-1)
            if nl_coef_b.numel() == 1:        # a scalar tensor (e.g. a learnable coefficient): the same for every sample
                nl_coef_b = nl_coef_b.expand(batch)
            nl_coef_b = nl_coef_b.contiguous()
            if nl_coef_b.numel() != batch:
                raise ValueError("a batched coefficient must have one entry per sample")
            self._keep.append(nl_coef_b)
            desc.nl_coef_b = nl_coef_b.data_ptr()
        dk, dkraw = f_mesh.wavenumber_tables()
        for i in range(self.n_dim):
            self._keep += [dk[i], dkraw[i]]
            desc.dk[i] = dk[i].data_ptr()
            desc.dkraw[i] = dkraw[i].data_ptr()
        # ---- coefficient tables (reference expressions, then re-laid out)
        self.tables_full = {}
        tab_channels = 1
        if tables is None:
            L = linear_coef
            if L is None:                                    # operator/_base.py:473-478
                L = torch.tensor([0.0], dtype=self.cdtype, device=self.device).reshape([1] * (self.n_dim + 2))
            tables = build_tables(integrator, dt, L, **integrator_cfg)
        if integrator in ("ETDRK2", "SETDRK2") and "coef_3" not in tables:
            # derived table for the one-read-less form of the two-stage step (include/fsm_b200.h, tab_coef)
            tables = dict(tables)
            tables["coef_3"] = tables["coef_1"] - tables["coef_2"]
        self.tables_full = tables
        # odd-order linear terms (KdV dispersion, advection) make exp(L dt) complex: 1-D kernels take complex tables
        self.complex_tables = self.n_dim == 1 and (has_imag(linear_coef) or any(has_imag(t) for t in tables.values()))
        if self.complex_tables and any(int(k) >= n // 2 for k, n in zip(kmax, self.shape)) and program != _cabi.PROG_LINEAR:
            raise NotImplementedError("complex linear coefficients need a dealiasing rate below 1 on the fused path "
                                      "(the Nyquist mode of the half spectrum is not Hermitian under a complex symbol)")
        desc.tab_complex = 1 if self.complex_tables else 0
        rot = {}
        tab_batch = 1       # tensor-valued coefficients of linear terms make every table per-sample (_base.py:339-357)
        for k, t in tables.items():
            rt = real_table(t)
            rot[k] = rt
            tab_batch, tab_channels = max(tab_batch, rt.shape[0]), max(tab_channels, rt.shape[1])
        if linear_coef is not None:
            rot["lin"] = real_table(linear_coef)
            tab_batch, tab_channels = max(tab_batch, rot["lin"].shape[0]), max(tab_channels, rot["lin"].shape[1])
        for k in list(rot):
            t = rot[k].expand(tab_batch, tab_channels, *self.shape).reshape(tab_batch * tab_channels, *self.shape)
            rot[k] = self._local_slab(_rot_half(t, self.shape))
            self._keep.append(rot[k])
        desc.tab_channels = tab_channels
        desc.tab_batched = 1 if tab_batch > 1 else 0
        self.tab_batch = tab_batch
        if "exp" in rot:
            desc.tab_exp = rot["exp"].data_ptr()
        if "half_exp" in rot:
            desc.tab_half_exp = rot["half_exp"].data_ptr()
        for i in range(6):
            if f"coef_{i + 1}" in rot:
                desc.tab_coef[i] = rot[f"coef_{i + 1}"].data_ptr()
        if "lin" in rot:
            desc.tab_lin = rot["lin"].data_ptr()
        def const_spectrum(t, what):
            t = _expand_table(t, self.shape)
            if t.shape[0] != 1:
                raise NotImplementedError(f"a per-sample {what} is not supported by the fused CUDA path")
            if t.shape[1] not in (1, n_channel):
                raise ValueError(f"{what} has an incompatible channel count")
            t = t[0].expand(n_channel, *self.shape).to(self.cdtype)
            t = self._local_slab(_rot_half(t, self.shape))
            self._keep.append(t)
            return t

        if source_hat is not None:
            self.source_rot = const_spectrum(source_hat, "explicit source")
            desc.source_hat = self.source_rot.data_ptr()
        if force_hat is not None:
            self.force_rot = const_spectrum(force_hat, "external force")
            desc.force_hat = self.force_rot.data_ptr()
        if self.P > 1:
            desc.slab_rank, desc.slab_nranks = self.rank, self.P
        self.rot_tables = rot
        self._desc = desc
        plan = ctypes.c_void_p()
        _cabi.check(lib.fsm_plan_create(ctypes.byref(plan), ctypes.byref(desc)), "plan_create")
        self._plan = plan
        self._lib = lib
        ws = lib.fsm_workspace_bytes(plan)
        self.ws_bytes = ws
        if not allocate:          # validation only (reference_adapter.lower): the plan exists, nothing can run
            self.workspace = None
            return
        self.workspace = torch.empty(max(ws, 16), dtype=torch.uint8, device=self.device)
        if self.P > 1:
            self._slab_counts = {}
            n1 = n2 = 0
            for op in range(4):
                a, b, c = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int32()
                _cabi.check(lib.fsm_slab_info(plan, op, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)), "slab_info")
                self._slab_counts[op] = (a.value, b.value)
                self.n_stages = c.value
                n1, n2 = max(n1, a.value), max(n2, b.value)
            # exchange buffers: kernels write/read them directly in rank-blocked layouts
            self._peer = slab[4] if len(slab) > 4 else None
            self._exch_mode = (slab[5] if len(slab) > 5 else None) or ("store" if self._peer is not None else "nccl")
            if self._exch_mode != "nccl" and self._peer is None:
                raise ValueError("the 'store' and 'dma' exchanges need a peer provider (torchfsm_b200.peer)")
            if self._peer is not None:
                self._recv, self._peer_idx = [], []
                for which, n in enumerate((n1, n2)):
                    t, ptrs, idx = self._peer.alloc(n, self.cdtype, self.device)
                    self._peer_idx.append(idx)
                    if len(ptrs) != self.P:
                        raise ValueError("peer provider must return one address per rank")
                    if self._exch_mode == "store":
                        # the kernels store into the destination ranks' receive buffers themselves
                        arr = (ctypes.c_void_p * self.P)(*[int(x) for x in ptrs])
                        _cabi.check(lib.fsm_slab_peers(plan, which + 1, arr, self.P), "slab_peers")
                    self._recv.append(t)
                if self._exch_mode == "store":
                    self._send = self._recv      # never written: the library ignores the send pointer in this mode
                else:                            # "dma": local send buffers, blocks pushed by the copy engines
                    self._send = [torch.empty(n, dtype=self.cdtype, device=self.device) for n in (n1, n2)]
                    self._remote = [[self._peer.remote(self._peer_idx[which], r) for r in range(self.P)] for which in range(2)]
            else:
                self._send = [torch.empty(n, dtype=self.cdtype, device=self.device) for n in (n1, n2)]
                self._recv = [torch.empty(n, dtype=self.cdtype, device=self.device) for n in (n1, n2)]
            nsub = int(slab[3]) if len(slab) > 3 and slab[3] else 0
            if self._exch_mode == "store":
                nsub = 1
            if nsub <= 0:       # sub-slabs pipeline the exchanges with the local y/z chain (profiles/r2_slab_scaling.md):
                # four where a sub-slab still fills the GPU for several waves, two on thinner slabs (measured at 512^3:
                # 4 ranks 8.82 ms with 4 sub-slabs against 9.17 with 2; 8 ranks 5.33 ms with 2 against 5.87 with 4)
                nsub = 4 if self.nxl >= 128 else (2 if self.nxl >= 16 else 1)
            while nsub > 1 and self.nxl % nsub:
                nsub //= 2
            self.nsub = max(1, nsub)
            self._comm_stream = torch.cuda.Stream(self.device) if self.device.type == "cuda" else None
            n_copy = int(slab[6]) if len(slab) > 6 and slab[6] else 1
            self._copy_streams = [torch.cuda.Stream(self.device) for _ in range(n_copy)] \
                if (n_copy > 1 and self.device.type == "cuda") else []
            # the whole step (every phase launch, copy, barrier and stream dependency of its stages) is captured in
            # ONE CUDA graph per state buffer and replayed: the phase loop is a few hundred host calls per step
            self._graph_enabled = bool(slab[7]) if len(slab) > 7 else True
            self._graphs, self._warm, self._graph_error = {}, False, None

    def _local_slab(self, t):
        """(X, n1, nh, n0) rot-half table -> contiguous local ky lines when the grid is slab-decomposed. Ownership is
        cyclic (rank r holds ky = r, r + P, ...), so every rank owns the same share of the dealiased band and the
        inverse-side exchange ships kept lines only (include/fsm_b200.h)."""
        if self.P == 1:
            return t
        return t[:, self.rank::self.P].contiguous()

    # ---- slab-decomposed phases ---------------------------------------------------------------------
    def _exchange(self, which, count, offset=0, async_op=False):
        import torch.distributed as dist
        if self._exch_mode == "store":      # the data already sits in the peers' buffers: separate writers and readers
            self._peer.barrier(self._peer_idx[which])
            return None
        if self._exch_mode == "dma":
            # block q of the send range goes to rank q's receive range, slot `rank`: P-1 peer copies on the copy
            # engines (+ one local), then a device-side barrier; nothing here occupies an SM for long
            blk = count // self.P
            src = self._send[which]
            fan = self._copy_streams if self.device.type == "cuda" else []
            if fan:     # copy_streams > 1: the block copies of one exchange spread over several streams
                cur = torch.cuda.current_stream(self.device)
                start = torch.cuda.Event()
                start.record(cur)
            for i in range(self.P):
                q = (self.rank + i) % self.P                 # stagger the destinations across ranks
                dst = self._remote[which][q][offset + self.rank * blk: offset + (self.rank + 1) * blk]
                if fan:
                    side = fan[i % len(fan)]
                    side.wait_event(start)
                    with torch.cuda.stream(side):
                        dst.copy_(src[offset + q * blk: offset + (q + 1) * blk], non_blocking=True)
                else:
                    dst.copy_(src[offset + q * blk: offset + (q + 1) * blk], non_blocking=True)
            for side in fan[:self.P]:
                done = torch.cuda.Event()
                done.record(side)
                cur.wait_event(done)
            self._peer.barrier(self._peer_idx[which])
            return _StreamWork(self.device) if async_op else None
        return dist.all_to_all_single(self._recv[which][offset:offset + count], self._send[which][offset:offset + count],
                                      group=self.group, async_op=async_op)

    def _slab_phase(self, op, stage, phase, u_hat, aux, which_send, which_recv, sub=0, nsub=1):
        snd = self._send[which_send].data_ptr() if which_send is not None else None
        rcv = self._recv[which_recv].data_ptr() if which_recv is not None else None
        _cabi.check(self._lib.fsm_slab_phase(self._plan, op, stage, phase, sub, nsub,
                                             u_hat.data_ptr() if u_hat is not None else None,
                                             aux.data_ptr() if aux is not None else None, self.workspace.data_ptr(),
                                             self.ws_bytes, snd, rcv, self._stream()), "slab_phase")

    def _slab_begin(self):
        """Direct exchange: the previous operation's last reads of the receive buffers must be over on every rank
        before this one's first remote stores."""
        if self.P > 1 and self._peer is not None:
            self._peer.barrier(self._peer_idx[0], channel=1)

    def _slab_eval(self, op, stage, u_hat, aux=None):
        """One nonlinear evaluation + stage combine on a slab-decomposed grid. With nsub > 1 sub-slabs the
        all-to-all of sub-slab h+1 (side stream) overlaps the local y/z chain of sub-slab h."""
        c1, c2 = self._slab_counts[op]
        H = self.nsub
        h1, h2 = c1 // H, c2 // H
        self._slab_phase(op, stage, 0, u_hat, aux, 0, None, 0, H)
        if H == 1 or self.device.type != "cuda":
            for h in range(H):
                self._exchange(0, h1, h * h1)
                self._slab_phase(op, stage, 1, u_hat, aux, 1, 0, h, H)
                self._exchange(1, h2, h * h2)
        else:
            cur = torch.cuda.current_stream(self.device)
            comm = self._comm_stream
            ev = torch.cuda.Event()
            ev.record(cur)
            with torch.cuda.stream(comm):
                comm.wait_event(ev)
                works1 = [self._exchange(0, h1, h * h1, async_op=True) for h in range(H)]
            works2 = []
            for h in range(H):
                works1[h].wait()                         # the compute stream waits for this sub-slab only
                self._slab_phase(op, stage, 1, u_hat, aux, 1, 0, h, H)
                evc = torch.cuda.Event()
                evc.record(cur)
                with torch.cuda.stream(comm):
                    comm.wait_event(evc)
                    works2.append(self._exchange(1, h2, h * h2, async_op=True))
            for w in works2:
                w.wait()
        self._slab_phase(op, stage, 2, u_hat, aux, None, 1, 0, H)

    # ---- plumbing ---------------------------------------------------------------------------
    def __del__(self):
        try:
            if getattr(self, "_plan", None):
                self._lib.fsm_plan_destroy(self._plan)
                self._plan = None
        except Exception:
            pass

    def _stream(self):
        if self.device.type == "cuda":
            return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        return ctypes.c_void_p(0)

    def info(self):
        a, b, c, d = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int32()
        _cabi.check(self._lib.fsm_plan_info(self._plan, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c),
                                            ctypes.byref(d)), "plan_info")
        return {"launches_per_step": a.value, "algo_bytes_per_step": b.value, "modes_per_field": c.value,
                "chunk": d.value}

    def touched_bytes(self):
        """Bytes each pass class really reads + writes per step (kept modes only), ``fsm_plan_traffic``."""
        by = (ctypes.c_int64 * 4)()
        _cabi.check(self._lib.fsm_plan_traffic(self._plan, by), "plan_traffic")
        return dict(zip(("IX", "MID", "PHYS", "FX"), [int(x) for x in by]))

    def stage_kinds(self):
        """Per integrator stage: index of the compile-time combine structure used by the forward-x epilogue
        (-1 = generic data-driven path)."""
        buf = (ctypes.c_int32 * 8)()
        n = self._lib.fsm_stage_kinds(self._plan, buf, 8)
        return [int(buf[i]) for i in range(min(n, 8))]

    def profile(self, on: bool):
        _cabi.check(self._lib.fsm_profile_enable(self._plan, 1 if on else 0), "profile_enable")

    def profile_read(self):
        ms, n, by = (ctypes.c_double * 4)(), (ctypes.c_int64 * 4)(), (ctypes.c_int64 * 4)()
        _cabi.check(self._lib.fsm_profile_read(self._plan, ms, n, by), "profile_read")
        names = ("IX", "MID", "PHYS", "FX")
        return {k: {"ms": ms[i], "launches": n[i], "algo_bytes_per_step": by[i]} for i, k in enumerate(names)}

    def empty_half(self):
        return torch.empty((self.B, self.C, self.nmodes), dtype=self.cdtype, device=self.device)

    # ---- native entry points (rot-half state) -------------------------------------------------
    def r2c(self, u: torch.Tensor) -> torch.Tensor:
        u = u.to(self.rdtype).contiguous()
        out = self.empty_half()
        if self.P > 1:
            _, c2 = self._slab_counts[2]
            self._slab_begin()
            self._slab_phase(2, 0, 1, out, u, 1, None)
            self._exchange(1, c2)
            self._slab_phase(2, 0, 2, out, u, None, 1)
            return out
        _cabi.check(self._lib.fsm_r2c(self._plan, u.data_ptr(), out.data_ptr(), self.workspace.data_ptr(),
                                      self.ws_bytes, self._stream()), "r2c")
        return out

    def c2r(self, u_hat: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if out is None:
            out = torch.empty((self.B, self.C) + self.local_shape, dtype=self.rdtype, device=self.device)
        elif (tuple(out.shape) != (self.B, self.C) + tuple(self.local_shape) or out.dtype != self.rdtype
              or not out.is_contiguous() or not _same_device(out.device, self.device)):
            raise ValueError("c2r(out=...) needs a contiguous (B, C, *grid) tensor of the plan's dtype and device")
        if self.P > 1:
            c1, _ = self._slab_counts[3]
            self._slab_begin()
            self._slab_phase(3, 0, 0, u_hat, out, 0, None)
            self._exchange(0, c1)
            self._slab_phase(3, 0, 1, u_hat, out, None, 0)
            return out
        _cabi.check(self._lib.fsm_c2r(self._plan, u_hat.data_ptr(), out.data_ptr(), self.workspace.data_ptr(),
                                      self.ws_bytes, self._stream()), "c2r")
        return out

    def spectral_map(self, u_hat: torch.Tensor, c_out: int, terms, dealias: bool = False) -> torch.Tensor:
        """out[b][co] = sum of coef * prod_a (i k_a)^p_a * (1/lap)^q * u_hat[b][ci] over ``terms`` =
        [(co, ci, (p0, p1, p2), q, coef)]; one point-wise kernel on the rot-half state (``fsm_spectral_map``)."""
        arr = (_cabi.FsmMapTerm * max(1, len(terms)))()
        for i, (co, ci, pw, q, coef) in enumerate(terms):
            arr[i].out_channel, arr[i].in_channel, arr[i].inv_laplacian, arr[i].coef = int(co), int(ci), int(q), float(coef)
            for a in range(3):
                arr[i].power[a] = int(pw[a]) if a < len(pw) else 0
        c_in = u_hat.shape[1]
        out = torch.empty((self.B, c_out, self.nmodes), dtype=self.cdtype, device=self.device)
        _cabi.check(self._lib.fsm_spectral_map(self._plan, u_hat.data_ptr(), c_in, out.data_ptr(), c_out, arr, len(terms),
                                               1 if dealias else 0, self._stream()), "spectral_map")
        return out

    def half_to_full(self, u_hat: torch.Tensor) -> torch.Tensor:
        if self.P > 1:
            raise NotImplementedError("full-spectrum frames are not available for slab-decomposed grids")
        out = torch.empty((self.B, self.C) + self.shape, dtype=self.cdtype, device=self.device)
        _cabi.check(self._lib.fsm_half_to_full(self._plan, u_hat.data_ptr(), out.data_ptr(), self._stream()),
                    "half_to_full")
        return out

    def full_to_half(self, full_hat: torch.Tensor) -> torch.Tensor:
        if self.P > 1:
            raise NotImplementedError("full-spectrum input is not available for slab-decomposed grids")
        full_hat = full_hat.to(self.cdtype).contiguous()
        out = self.empty_half()
        _cabi.check(self._lib.fsm_full_to_half(self._plan, full_hat.data_ptr(), out.data_ptr(), self._stream()),
                    "full_to_half")
        return out

    def step_half(self, u_hat: torch.Tensor, n_steps: int = 1) -> torch.Tensor:
        """Advance the rot-half state in place by ``n_steps``."""
        if getattr(self, "ks_group", None) is not None and self._desc.program == _cabi.PROG_KS \
                and self._desc.ks_remove_mean and int(n_steps) > 0:
            return self._step_half_ks_sharded(u_hat, int(n_steps))
        if self.P > 1 and self.n_stages and self._desc.program != _cabi.PROG_LINEAR:
            self._slab_begin()
            self._slab_steps(u_hat, int(n_steps))
            return u_hat
        _cabi.check(self._lib.fsm_step(self._plan, u_hat.data_ptr(), self.workspace.data_ptr(), self.ws_bytes,
                                       int(n_steps), self._stream()), "step")
        return u_hat

    def _slab_step_once(self, u_hat):
        for stage in range(self.n_stages):
            self._slab_eval(0, stage, u_hat)

    def _slab_steps(self, u_hat, n):
        """``n`` steps of a slab-decomposed grid. On CUDA the first step ever runs eagerly (it also warms every lazy
        initialisation: NCCL channels, peer mappings), then one step is captured into a CUDA graph keyed by the state
        buffer and replayed; every rank takes the same decisions, so collectives inside the graphs stay matched."""
        if self.device.type != "cuda" or not self._graph_enabled:
            for _ in range(n):
                self._slab_step_once(u_hat)
            return
        if not self._warm and n > 0:
            self._slab_step_once(u_hat)
            self._warm = True
            n -= 1
        if n <= 0:
            return
        key = (u_hat.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream)
        g = self._graphs.get(key)
        if g is None:
            try:
                g = torch.cuda.CUDAGraph()
                side = torch.cuda.Stream(self.device)
                side.wait_stream(torch.cuda.current_stream(self.device))
                with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local"):
                    self._slab_step_once(u_hat)
                torch.cuda.current_stream(self.device).wait_stream(side)
                self._graphs[key] = g
            except Exception as exc:      # capture unavailable: keep the eager phase loop
                self._graph_enabled, self._graph_error, g = False, repr(exc), None
        for _ in range(n):
            if g is not None:
                g.replay()
            else:
                self._slab_step_once(u_hat)

    _KS_FINAL_WEIGHTS = {"ETDRK1": [("coef_1", 1.0)], "SETDRK1": [("coef_1", 1.0)],
                         "ETDRK2": [("coef_1", 1.0, "coef_2", -1.0), ("coef_2", 1.0)],
                         "SETDRK2": [("coef_1", 1.0, "coef_2", -1.0), ("coef_2", 1.0)],
                         "SETDRK3": [("coef_3", 1.0), ("coef_4", 1.0), ("coef_5", 1.0)],
                         "SETDRK4": [("coef_4", 1.0), ("coef_5", 2.0), ("coef_5", 2.0), ("coef_6", 1.0)]}

    def _step_half_ks_sharded(self, u_hat: torch.Tensor, n_steps: int) -> torch.Tensor:
        """KS ensemble sharded over the ranks of ``ks_group``: the batch mean of _ks_convection.py:34-36 spans
        all ranks but only touches the k=0 bin, which never feeds back. Every rank steps with its local mean
        while the library logs the per-evaluation local zero-mode sums; ONE all-reduce of that log afterwards
        gives the exact correction of the zero mode (no collective inside the step)."""
        import torch.distributed as dist
        if getattr(self, "tab_batch", 1) > 1:
            raise NotImplementedError("sharded KS ensembles do not take per-sample coefficients")
        if self.integrator not in self._KS_FINAL_WEIGHTS:
            raise NotImplementedError(f"sharded KS ensembles are not supported with the {self.integrator} integrator")
        spec = self._KS_FINAL_WEIGHTS[self.integrator]
        n_stage = len(spec)
        log = torch.zeros(n_steps * n_stage, dtype=self.rdtype, device=self.device)
        _cabi.check(self._lib.fsm_ks_log(self._plan, log.data_ptr(), log.numel()), "ks_log")
        try:
            _cabi.check(self._lib.fsm_step(self._plan, u_hat.data_ptr(), self.workspace.data_ptr(), self.ws_bytes,
                                           n_steps, self._stream()), "step")
        finally:
            _cabi.check(self._lib.fsm_ks_log(self._plan, None, 0), "ks_log")
        glob = torch.cat([log, torch.tensor([float(self.B)], dtype=self.rdtype, device=self.device)])
        dist.all_reduce(glob, group=self.ks_group)
        delta = (log / self.B - glob[:-1] / glob[-1]).reshape(n_steps, n_stage)     # local mean - global mean
        tab0 = {k: v.reshape(v.shape[0], -1)[0, 0] for k, v in self.rot_tables.items()}
        w = []
        for term in spec:
            x = tab0[term[0]] * term[1]
            if len(term) == 4:
                x = x + tab0[term[2]] * term[3]
            w.append(x)
        w = torch.stack(w)
        e0 = tab0["exp"]
        powers = e0 ** torch.arange(n_steps - 1, -1, -1, device=self.device, dtype=self.rdtype)
        corr = (powers[:, None] * w[None, :] * delta).sum()
        u_hat[:, 0, 0] += corr
        return u_hat

    def rhs_half(self, u_hat: torch.Tensor) -> torch.Tensor:
        out = self.empty_half()
        if self.P > 1 and self._desc.program != _cabi.PROG_LINEAR:
            self._slab_begin()
            self._slab_eval(1, 0, u_hat, out)
            return out
        _cabi.check(self._lib.fsm_rhs(self._plan, u_hat.data_ptr(), out.data_ptr(), self.workspace.data_ptr(),
                                      self.ws_bytes, self._stream()), "rhs")
        return out

    # ---- reference integrator protocol (full spectra) -------------------------------------------
    def step(self, u_hat_full: torch.Tensor) -> torch.Tensor:
        h = self.full_to_half(u_hat_full)
        return self.half_to_full(self.step_half(h, 1))

    def forward(self, u_hat_full: torch.Tensor, dt: float) -> torch.Tensor:
        return self.step(u_hat_full)


class _StageLoopStepper(FusedStepper):
    """A stepper whose integrator stages are driven one by one from the host: between two stages the caller evaluates
    something on the stage state with the library's own passes (``_evaluate``) and hands it to ``_run_stage``."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        if self.P > 1:
            raise NotImplementedError("host-driven stages are not available on slab-decomposed grids")
        self._stage_off = []
        off = ctypes.c_int64()
        n = self._lib.fsm_stage_input(self._plan, -1, ctypes.byref(off))
        if n < 0:
            _cabi.check(n, "stage_input")
        for s in range(n):
            _cabi.check(min(0, self._lib.fsm_stage_input(self._plan, s, ctypes.byref(off))), "stage_input")
            self._stage_off.append(off.value)
        self._state_bytes = self.B * self.C * self.nmodes * (8 if self.rdtype == torch.float32 else 16)

    def _stage_state(self, s, u_hat):
        off = self._stage_off[s]
        if off < 0:
            return u_hat
        return self.workspace[off:off + self._state_bytes].view(self.cdtype).view(self.B, self.C, self.nmodes)

    def mask_state(self, x_hat: torch.Tensor) -> torch.Tensor:
        """Zero the modes outside the dealiasing box, in place (operator/_base.py:381-385)."""
        _cabi.check(self._lib.fsm_mask_state(self._plan, x_hat.data_ptr(), x_hat.shape[1], self._stream()), "mask_state")
        return x_hat

    def sym_outer(self, u: torch.Tensor) -> torch.Tensor:
        C = u.shape[1]
        out = torch.empty((u.shape[0], C * (C + 1) // 2) + tuple(u.shape[2:]), dtype=u.dtype, device=u.device)
        _cabi.check(self._lib.fsm_sym_outer(self._plan, u.data_ptr(), out.data_ptr(), C, self._stream()), "sym_outer")
        return out

    def step_half(self, u_hat: torch.Tensor, n_steps: int = 1) -> torch.Tensor:
        for _ in range(int(n_steps)):
            for s in range(len(self._stage_off)):
                self._run_stage(s, u_hat, self._evaluate(self._stage_state(s, u_hat)), None)
        return u_hat

    def rhs_half(self, u_hat: torch.Tensor) -> torch.Tensor:
        out = self.empty_half()
        self._run_stage(-1, u_hat, self._evaluate(u_hat), out)
        return out


class HostComposedStepper(_StageLoopStepper):
    """Integrator for nonlinear cores without a fused program (``_EXTERNAL_KINDS``): the plan carries the linear part
    and the stage formulas; each stage's nonlinear term is evaluated by ``nonlinear(stepper, stage_state_hat)`` from
    the library's own passes and combined by ``fsm_stage_combine`` (same tables, same stage algebra as the fused
    path; integrator/_etdrk.py:47-82, _setdrk_step.py:5-82, _rk.py:43-58)."""

    def __init__(self, *args, nonlinear=None, **kwargs):
        super().__init__(*args, **kwargs)
        self._nonlinear = nonlinear

    def _evaluate(self, x_hat):
        return self._nonlinear(self, x_hat)

    def _run_stage(self, s, u_hat, fresh, out):
        _cabi.check(self._lib.fsm_stage_combine(self._plan, s, u_hat.data_ptr(), fresh.data_ptr(),
                                                out.data_ptr() if out is not None else None,
                                                self.workspace.data_ptr(), self.ws_bytes, self._stream()), "stage_combine")


class DynamicForceStepper(_StageLoopStepper):
    """``NSPressureConvection(external_force)`` with a force operator that depends on the state
    (dedicated/_navier_stokes.py:237-254): before every stage the force is evaluated on the stage state (un-dealiased,
    as the reference does) by the force operator's own lowering -- any operator this package can evaluate -- and the
    fused convection + projection + combine of the stage runs with it (``fsm_stage_run``)."""

    def __init__(self, *args, force=None, **kwargs):
        super().__init__(*args, dynamic_force=True, **kwargs)
        self._force = force

    def _evaluate(self, x_hat):
        f_hat, c = self._force._eval_half(x_hat, self.f_mesh, self.C)
        if c != self.C:
            if c != 1:
                raise ValueError("the external force must have one channel or as many as the velocity")
            f_hat = f_hat.expand(self.B, self.C, self.nmodes)
        return f_hat.contiguous()

    def _run_stage(self, s, u_hat, f_hat, out):
        _cabi.check(self._lib.fsm_stage_run(self._plan, s, u_hat.data_ptr(), f_hat.data_ptr(),
                                            out.data_ptr() if out is not None else None,
                                            self.workspace.data_ptr(), self.ws_bytes, self._stream()), "stage_run")


class ExplicitRKStepper:
    """The explicit Runge-Kutta family other than RK4 (integrator/_rk.py:43-58, 82-255; non-adaptive): every k_i is one
    right-hand-side evaluation of the wrapped plan (``fsm_rhs``: the fused program, or the host-composed / forced
    variants), the stage states and the update are ``fsm_lincomb`` launches. Everything else (transforms, layout
    converters, introspection) is the wrapped stepper's."""

    def __init__(self, inner: FusedStepper, name: str, dt: float):
        self._inner, self.integrator, self.dt = inner, name, dt
        self._rows, self._b = RK_TABLEAUS[name]

    def __getattr__(self, item):
        return getattr(self._inner, item)

    def _lincomb(self, base, coefs, ks):
        terms = [(c, k) for c, k in zip(coefs, ks) if c != 0]
        out = torch.empty_like(base)
        n = len(terms)
        ptrs = (ctypes.c_void_p * max(1, n))(*[k.data_ptr() for _, k in terms])
        cf = (ctypes.c_double * max(1, n))(*[float(c) for c, _ in terms])
        st = self._inner
        _cabi.check(st._lib.fsm_lincomb(st._plan, out.data_ptr(), base.data_ptr(), ptrs, cf, n, base.numel(), st._stream()),
                    "lincomb")
        return out

    def step_half(self, u_hat: torch.Tensor, n_steps: int = 1) -> torch.Tensor:
        st, dt = self._inner, self.dt
        x = u_hat
        for _ in range(int(n_steps)):
            ks = [st.rhs_half(x)]
            for row in self._rows:                                  # _rk.py:50-52
                ks.append(st.rhs_half(self._lincomb(x, [dt * a for a in row[1:]], ks)))
            x = self._lincomb(x, [dt * b for b in self._b], ks)    # _rk.py:53 (zip stops at the shorter list)
        if x is not u_hat:
            u_hat.copy_(x)
        return u_hat

    def step(self, u_hat_full: torch.Tensor) -> torch.Tensor:
        st = self._inner
        return st.half_to_full(self.step_half(st.full_to_half(u_hat_full), 1))

    def forward(self, u_hat_full: torch.Tensor, dt: float) -> torch.Tensor:
        return self.step(u_hat_full)


# ------------------------------------------------------------------------------------------------
# Operators
# ------------------------------------------------------------------------------------------------
class _LinearFn(torch.autograd.Function):
    """y = A u for a real-linear A evaluated by the library (no graph inside); backward applies A^T, which for
    A = ifft . S . fft on real fields is ifft . S^H . fft: the same kernels with the conjugate-transposed symbol."""

    @staticmethod
    def forward(ctx, u, fwd, bwd):
        ctx.bwd = bwd
        return fwd(u.detach())

    @staticmethod
    def backward(ctx, g):
        return ctx.bwd(g.detach().contiguous()), None, None


class LinearCoef:
    """User-defined linear symbol (the reference's core protocol, operator/_base.py:16-54): ``__call__(f_mesh, n_channel)``
    returns L(k) in the reference layout ``(1|B, C, N...)``, built from the tables of ``FourierMesh``."""

    def __call__(self, f_mesh: FourierMesh, n_channel: int) -> torch.Tensor:
        raise NotImplementedError

    def nonlinear_like(self, u_fft, f_mesh, u=None):
        return self(f_mesh, u_fft.shape[1]) * u_fft


class NonlinearFunc:
    """User-defined nonlinear core (operator/_base.py:56-103): ``__call__(u_fft, f_mesh, u)`` takes the full spectrum
    ``(B, C, N...)`` (dealiased when ``dealiasing_swtich``) and the matching physical field and returns the full
    spectrum of the term. ``f_mesh.fft`` / ``f_mesh.ifft`` inside it run on the library's passes."""

    def __init__(self, dealiasing_swtich: bool = True) -> None:
        self._dealiasing_swtich = dealiasing_swtich

    def __call__(self, u_fft, f_mesh, u=None) -> torch.Tensor:
        raise NotImplementedError

    def spatial_value(self, u_fft, f_mesh, u=None):
        return f_mesh.ifft(self(u_fft, f_mesh, u)).real


class CoreGenerator:
    """``__call__(f_mesh, n_channel) -> LinearCoef | NonlinearFunc`` (operator/_base.py:106-126): decides the core once
    the mesh and the channel count are known."""

    def __call__(self, f_mesh: FourierMesh, n_channel: int):
        raise NotImplementedError


class OperatorLike:
    """Sum of generator terms (mirror of ``OperatorLike``/``Operator``, operator/_base.py:286-850).

    ``OperatorLike(operator_generators, coefs)`` as in the reference (:297-307): every generator is a ``LinearCoef``, a
    ``NonlinearFunc``, a ``CoreGenerator`` or a callable ``(f_mesh, n_channel) -> core``. The built-in operators pass
    ready-made terms instead."""

    def __init__(self, terms=None, coefs: Optional[List] = None):
        terms = [] if terms is None else (list(terms) if isinstance(terms, (list, tuple)) else [terms])
        coefs = list(coefs) if coefs is not None else [1] * len(terms)
        if len(coefs) != len(terms):
            raise ValueError("The length of coefs should match the number of operator generators")
        unit = lambda c: not isinstance(c, torch.Tensor) and c == 1      # noqa: E731
        self.terms: List[_Term] = [(t if unit(c) else t.scaled(c)) if isinstance(t, _Term) else
                                   _Term("custom", c, {"generator": t}) for t, c in zip(terms, coefs)]
        self._rt = None               # terms with the user-defined generators resolved for the registered mesh
        self._de_aliasing_rate = 2 / 3
        self._integrator = "auto"
        self._integrator_config = {}
        self._value_mesh_check_func: Callable[[int, int], bool] = lambda dim_value, dim_mesh: True
        self._state_dict = {"f_mesh": None, "n_channel": None, "linear_coef": None, "integrator": None}
        self._lowered = None
        self._chunk = 0
        self._lanes = 0
        self._slab = None

    # ---- algebra (operator/_base.py:170-206, 826-850) ------------------------------------------
    def _new(self, terms):
        op = Operator(terms)
        op._value_mesh_check_func = self._value_mesh_check_func
        return op

    def __add__(self, other):
        if isinstance(other, OperatorLike):
            return self._new(self.terms + other.terms)
        if isinstance(other, torch.Tensor):
            return self._new(self.terms + [_Term("explicit_source", 1, {"source": other})])
        return NotImplemented

    __radd__ = __add__
    __iadd__ = __add__

    def __mul__(self, other):
        if isinstance(other, OperatorLike):
            return NotImplemented
        return self._new([t.scaled(other) for t in self.terms])

    __rmul__ = __mul__
    __imul__ = __mul__

    def __neg__(self):
        return self._new([t.scaled(-1) for t in self.terms])

    def __sub__(self, other):
        try:
            return self + (-1 * other)
        except Exception:
            return NotImplemented

    def __rsub__(self, other):
        try:
            return other + (-1 * self)
        except Exception:
            return NotImplemented

    def __truediv__(self, other):
        try:
            return self * (1 / other)
        except Exception:
            return NotImplemented

    # ---- configuration ------------------------------------------------------------------------------
    def set_integrator(self, integrator, **integrator_config):
        """operator/_base.py:646-674"""
        if isinstance(integrator, str):
            assert integrator == "auto", ("The integrator should be 'auto' or an instance of ETDRKIntegrator, "
                                          "SETDRKIntegrator or RKIntegrator")
        else:
            assert isinstance(integrator, (ETDRKIntegrator, SETDRKIntegrator, RKIntegrator)), (
                "The integrator should be 'auto' or an instance of ETDRKIntegrator, SETDRKIntegrator or RKIntegrator")
        self._integrator = integrator
        self._integrator_config = integrator_config
        self._state_dict["integrator"] = None

    def set_de_aliasing_rate(self, de_aliasing_rate: float):
        """operator/_base.py:273-281 (the reference's version breaks the next call; this one re-registers)."""
        self._de_aliasing_rate = de_aliasing_rate
        self._state_dict["integrator"] = None
        self._lowered = None

    def set_slab_decomposition(self, group=None, rank: Optional[int] = None, nranks: Optional[int] = None,
                               nsub: int = 0, exchange: str = "nccl", peers=None, copy_streams: int = 1,
                               graph: bool = True):
        """Decompose ONE 3-D grid over the ranks of a torch.distributed process group:
        ``integrate`` / ``__call__`` then take and return the local physical x-slab
        ``(B, C, n0/P, n1, n2)`` of the rank. The two transposes per evaluation run as ``exchange`` =

        * ``"nccl"``: all-to-alls of rank-blocked send buffers;
        * ``"dma"``: the same send buffers, each block pushed into its owner's receive buffer by the copy
          engines (peer-to-peer copies over NVLink), so no SM is taken from the transforms;
        * ``"store"``: no send buffer, the transform kernels store straight into the destination ranks'
          receive buffers; an exchange is only a device-side barrier.

        ``"dma"`` and ``"store"`` need peer-visible buffers: ``peers`` is a provider from ``torchfsm_b200.peer``
        (default: torch symmetric memory). ``nsub`` sub-slabs (0 = choose) pipeline the exchange of one part
        with the local work on another ("nccl", "dma"); ``copy_streams`` spreads the block copies of one "dma"
        exchange over several streams; ``graph`` captures each step in one CUDA graph (replayed per step).

        Spectral ownership is cyclic in ky (rank r holds ky = r, r + P, ...): every rank owns the same share of the
        dealiased band and the inverse-side exchange ships kept lines only."""
        import torch.distributed as dist
        if exchange not in ("nccl", "dma", "store"):
            raise ValueError(f"unknown exchange {exchange!r}")
        if exchange != "nccl" and peers is None:
            from .peer import SymmetricMemoryPeers
            peers = SymmetricMemoryPeers(group)
        self._slab = (dist.get_rank(group) if rank is None else rank,
                      dist.get_world_size(group) if nranks is None else nranks, group, nsub,
                      peers if exchange != "nccl" else None, exchange, int(copy_streams), bool(graph))
        self._state_dict["integrator"] = None
        self._rhs_stepper = None

    def set_ensemble_group(self, group):
        """The batch is sharded over the ranks of ``group`` (one process per GPU). Samples are independent
        except for KSConvection(remove_mean=True), whose mean spans every rank; see
        ``FusedStepper._step_half_ks_sharded``."""
        self._ensemble_group = group
        if self._state_dict.get("integrator") is not None:
            self._state_dict["integrator"].ks_group = group

    def to(self, device=None, dtype=None):
        """operator/_base.py:792-803: move the registered mesh (and with it every table) to another device / dtype;
        the integrator is rebuilt on the next call."""
        sd = self._state_dict
        if sd is not None and sd.get("f_mesh") is not None:
            self.register_mesh(FourierMesh(sd["f_mesh"], device=device, dtype=dtype), sd["n_channel"])
        return self

    def set_chunk(self, chunk: int):
        """Samples per pass launch (0 = library default); a tuning knob of the CUDA path."""
        self._chunk = int(chunk)
        self._state_dict["integrator"] = None

    def set_lanes(self, lanes: int):
        """Ensembles: number of independent sample ranges the library steps concurrently on internal streams
        (0 = library default, 1 = everything on the caller's stream); a tuning knob of the CUDA path."""
        self._lanes = int(lanes)
        self._state_dict["integrator"] = None

    def set_gradient_checkpointing(self, enabled: bool = True):
        """Gradient mode (autograd.py) keeps every intermediate of every step, as the reference's autograd does; with
        checkpointing only the state at step boundaries is kept and each step is recomputed in backward."""
        self._grad_checkpoint = bool(enabled)
        return self

    def register_additional_check(self, func: Callable[[int, int], bool]):
        self._value_mesh_check_func = func

    @property
    def _terms_now(self) -> List[_Term]:
        return self._rt if self._rt is not None else self.terms

    def _resolve_terms(self, f_mesh: FourierMesh, n_channel: int) -> List[_Term]:
        """User-defined generators -> what they produce on this mesh (operator/_base.py:611-624): a ``LinearCoef``
        becomes a tabulated symbol, a ``NonlinearFunc`` a host-composed nonlinear term."""
        out = []
        for t in self.terms:
            if t.kind != "custom":
                out.append(t)
                continue
            core = t.params["generator"]
            if not isinstance(core, (LinearCoef, NonlinearFunc)):
                core = core(f_mesh, n_channel)
            if isinstance(core, LinearCoef) or (hasattr(core, "nonlinear_like") and not hasattr(core, "_dealiasing_swtich")):
                out.append(_Term("linear_tensor", t.coef, {"L": core(f_mesh, n_channel)}))
            elif isinstance(core, NonlinearFunc) or hasattr(core, "_dealiasing_swtich"):
                out.append(_Term("custom_nonlinear", t.coef, {"func": core}))
            else:
                raise TypeError("a core generator must produce a LinearCoef or a NonlinearFunc")
        return out

    @property
    def is_linear(self) -> bool:
        return all(t.kind in _LINEAR_KINDS for t in self._terms_now)

    # ---- lowering ----------------------------------------------------------------------------------
    def register_mesh(self, mesh, n_channel: int, device=None, dtype=None):
        """operator/_base.py:581-624: build L and classify the nonlinear part into a fused program."""
        f_mesh = mesh if isinstance(mesh, FourierMesh) and device is None and dtype is None \
            else FourierMesh(mesh, device=device, dtype=dtype)
        self._state_dict = {"f_mesh": f_mesh, "n_channel": n_channel, "linear_coef": None, "integrator": None}
        self._rhs_stepper = None
        self._rt = self._resolve_terms(f_mesh, n_channel) if any(t.kind == "custom" for t in self.terms) else None
        terms = self._terms_now
        kinds = {t.kind for t in terms}
        if kinds & set(_COMPOSITE_KINDS):
            if len(terms) != 1 or isinstance(terms[0].coef, torch.Tensor):
                raise NotImplementedError("Velocity2Pressure / Vorticity2Pressure cannot be summed with other operators "
                                          "on the fused CUDA path")
            t = terms[0]
            if t.kind == "vorticity2pressure" and (f_mesh.n_dim != 2 or n_channel != 1):
                raise ValueError("Only vorticity in 2Dmesh is supported")
            if t.kind == "velocity2pressure" and f_mesh.n_dim != n_channel:
                raise ValueError("convection operator only works for vector field with the same dimension as mesh")
            self._lowered = dict(composite=t, c_out=1)
            return
        if kinds & set(_MAP_KINDS):
            c_out, terms = self._lower_map(f_mesh, n_channel)
            self._lowered = dict(map=terms, c_out=c_out)
            return
        lin = []
        program, nl_coef, ks_remove_mean = _cabi.PROG_LINEAR, 0.0, True
        source_hat = force_hat = dyn_force = nl_coef_b = None
        external = []
        for t in terms:
            if t.kind in _LINEAR_KINDS:
                lin.append(t)
            elif t.kind == "ks_convection" and f_mesh.n_dim == 1:
                # no fused 1-D program for 1/2 (phi_x)^2: composed on the host from the library's passes
                if n_channel != 1:
                    raise NotImplementedError("KSConvection only supports scalar field")
                if isinstance(t.coef, torch.Tensor) and not _per_sample(t.coef):
                    raise NotImplementedError("tensor-valued coefficients on nonlinear terms must be per-sample scalars")
                external.append(t)
            elif t.kind in _PROGRAM_OF:
                if program != _cabi.PROG_LINEAR:
                    raise NotImplementedError("only one convective nonlinear term per operator is supported "
                                              "by the fused CUDA path")
                if isinstance(t.coef, torch.Tensor):          # one coefficient per sample (operator/_base.py:375-403)
                    if not _per_sample(t.coef):
                        raise NotImplementedError("tensor-valued coefficients on nonlinear terms must be per-sample scalars")
                    if t.kind == "ns_pressure_convection" and t.params.get("external_force") is not None:
                        raise NotImplementedError("a per-sample coefficient on NSPressureConvection cannot be combined "
                                                  "with an external force")
                    program, nl_coef, nl_coef_b = _PROGRAM_OF[t.kind], 1.0, t.coef
                else:
                    program, nl_coef = _PROGRAM_OF[t.kind], float(t.coef)
                if t.kind == "convection" and f_mesh.n_dim != n_channel:
                    raise ValueError("convection operator only works for vector field with the same dimension as mesh")
                if t.kind == "vorticity_convection" and (f_mesh.n_dim != 2 or n_channel != 1):
                    raise ValueError("Only vorticity in 2Dmesh is supported")
                if t.kind == "ks_convection":
                    if n_channel != 1:
                        raise NotImplementedError("KSConvection only supports scalar field")
                    ks_remove_mean = bool(t.params.get("remove_mean", True))
                if t.kind == "ns_pressure_convection":
                    if f_mesh.n_dim != n_channel or f_mesh.n_dim < 2:
                        raise ValueError("convection operator only works for vector field with the same dimension as mesh")
                    force = t.params.get("external_force")
                    if force is not None:
                        # _navier_stokes.py:237-254: coef * (grad lap^-1 div(conv - f) - (conv - f)); the fused epilogue
                        # adds -coef * f_hat to coef * conv_hat before the projection. The force is evaluated once:
                        # it must not depend on the state (explicit sources only).
                        if any(ft.kind != "explicit_source" for ft in force.terms):
                            # the force depends on the state: evaluated before every stage by its own lowering
                            dyn_force = force
                            continue
                        f_hat = None
                        for ft in force.terms:
                            fs = ft.params["source"].to(device=f_mesh.device)
                            fh = ft.coef * (fs if ft.params.get("in_fourier") else torch.fft.fftn(fs, dim=list(range(2, fs.dim()))))
                            f_hat = fh if f_hat is None else f_hat + fh
                        # the reference subtracts the force from the convection in place and then adds it once more
                        # (:241-254: `convection -= force` ... `- convection + force`): coef * (P(conv - f) + f) with
                        # P(c) = grad lap^-1 div c - c. The second copy rides on the constant-source slot, after the projection.
                        force_hat = -float(t.coef) * f_hat
                        source_hat = float(t.coef) * f_hat if source_hat is None else source_hat + float(t.coef) * f_hat
            elif t.kind in _EXTERNAL_KINDS:
                if isinstance(t.coef, torch.Tensor) and not _per_sample(t.coef):
                    raise NotImplementedError("tensor-valued coefficients on nonlinear terms must be per-sample scalars")
                if t.kind == "conservative_convection" and f_mesh.n_dim != n_channel:
                    raise ValueError("div operator only works for vector field with the same dimension as mesh")
                external.append(t)
            elif t.kind == "explicit_source":
                src = t.params["source"].to(device=f_mesh.device)
                s_hat = src if t.params.get("in_fourier") else torch.fft.fftn(src, dim=list(range(2, src.dim())))   # _base.py:1002-1005
                s_hat = t.coef * s_hat
                source_hat = s_hat if source_hat is None else source_hat + s_hat
            else:
                raise ValueError(f"Operator {t.kind} is not supported")
        L = None
        if lin:                                                                   # operator/_base.py:339-357
            L = sum(t.coef * self._linear_core(t, f_mesh, n_channel) for t in lin)
        if external and program != _cabi.PROG_LINEAR:
            raise NotImplementedError("ImplicitSource(func) / ConservativeConvection cannot be summed with a fused "
                                      "convective term on the CUDA path")
        kmax = f_mesh.low_pass_kmax(self._de_aliasing_rate) if (program != _cabi.PROG_LINEAR or external) \
            else [n // 2 for n in f_mesh.shape]
        if program == _cabi.PROG_NS3D and any(k >= n // 2 and n % 2 == 0 for k, n in zip(kmax, f_mesh.shape)):
            # the reference's pressure projection leaves anti-Hermitian content on the Nyquist planes of its full
            # C2C state, which feeds back through i*k_N in later evaluations (_navier_stokes.py:249-254); the
            # half-spectrum state cannot carry it, so results drift apart after the first step (1e-5 by step 2)
            raise NotImplementedError("NSPressureConvection needs a de-aliasing rate that removes the Nyquist planes "
                                      "(rate < 1) on the fused CUDA path")
        self._state_dict["linear_coef"] = L
        self._lowered = dict(program=program, nl_coef=nl_coef, ks_remove_mean=ks_remove_mean,
                             source_hat=source_hat, force_hat=force_hat, kmax=kmax, external=external, dyn_force=dyn_force,
                             nl_coef_b=nl_coef_b)

    def _lower_map(self, f_mesh: FourierMesh, n_channel: int):
        """Operators made of symbol products only (Grad, Div, Curl, Vorticity2Velocity, optionally summed with linear
        cores) -> (output channels, [(out, in, powers, inverse-Laplacian power, coefficient)]) for ``fsm_spectral_map``.
        generic/_grad.py:6-28, _div.py:9-36, _curl.py:9-75, dedicated/_navier_stokes.py:75-104."""
        d, C = f_mesh.n_dim, n_channel

        def e(a, order=1):
            return tuple(order if i == a else 0 for i in range(3))

        per_term = []
        for t in self._terms_now:
            if isinstance(t.coef, torch.Tensor):
                raise NotImplementedError("tensor-valued coefficients are not supported on channel-changing operators")
            k, c = t.kind, float(t.coef)
            if k == "grad":
                if C != 1:
                    raise ValueError("The Grad operator only supports scalar field.")
                per_term.append((d, [(a, 0, e(a), 0, c) for a in range(d)]))
            elif k == "div":
                if d != C:
                    raise ValueError("div operator only works for vector field with the same dimension as mesh")
                per_term.append((1, [(0, a, e(a), 0, c) for a in range(d)]))
            elif k == "curl":
                if d != C:
                    raise ValueError("div operator only works for vector field with the same dimension as mesh")
                if C > 3 or C < 2:
                    raise ValueError("div operator only works for 2D or 3D vector field")
                if C == 2:
                    per_term.append((1, [(0, 1, e(0), 0, c), (0, 0, e(1), 0, -c)]))
                else:
                    per_term.append((3, [(0, 2, e(1), 0, c), (0, 1, e(2), 0, -c), (1, 0, e(2), 0, c), (1, 2, e(0), 0, -c),
                                         (2, 1, e(0), 0, c), (2, 0, e(1), 0, -c)]))
            elif k == "vorticity2velocity":
                if d != 2 or C != 1:
                    raise ValueError("Only vorticity in 2Dmesh is supported")
                per_term.append((2, [(0, 0, e(1), 1, -c), (1, 0, e(0), 1, c)]))
            elif k == "laplacian":
                per_term.append((C, [(ch, ch, e(a, 2), 0, c) for ch in range(C) for a in range(d)]))
            elif k == "biharmonic":
                per_term.append((C, [(ch, ch, tuple(x + y for x, y in zip(e(a, 2), e(b, 2))), 0, c * (1 if a == b else 2))
                                     for ch in range(C) for a in range(d) for b in range(a, d)]))
            elif k == "spatial_derivative":
                if C != 1:
                    raise ValueError("The SpatialDerivative operator only supports scalar field.")
                per_term.append((1, [(0, 0, e(t.params["dim_index"], t.params["order"]), 0, c)]))
            elif k == "implicit_unit_source":
                per_term.append((C, [(ch, ch, (0, 0, 0), 0, c) for ch in range(C)]))
            elif k == "linear_tensor":
                raise NotImplementedError("a tabulated linear symbol has no point-wise map form")
            else:
                raise NotImplementedError(f"{k} cannot be summed with a channel-changing operator on the fused CUDA path")
        c_out = per_term[0][0]
        if any(co != c_out for co, _ in per_term):
            raise NotImplementedError("terms of one operator must produce the same number of channels on the fused CUDA path")
        terms = [x for _, lst in per_term for x in lst]
        if len(terms) > 32 or max(C, c_out) > 4:
            raise NotImplementedError("too many symbol terms / channels for one point-wise map")
        return c_out, terms

    def _tf(self, batch: int, n_channel: int) -> "FusedStepper":
        """Transform-only plan (r2c, c2r, layout converters, spectral map) for ``batch`` fields of ``n_channel`` channels."""
        f_mesh = self._state_dict["f_mesh"]
        cache = self.__dict__.setdefault("_tf_steppers", {})
        key = (id(f_mesh), batch, n_channel, self._de_aliasing_rate, id(self._slab))
        st = cache.get(key)
        if st is None:
            if len(cache) > 8:
                cache.clear()
            st = FusedStepper(f_mesh, batch, n_channel, _cabi.PROG_LINEAR, "RK4", 1.0, None, 0.0, None,
                              f_mesh.low_pass_kmax(self._de_aliasing_rate), True, {}, slab=self._slab)
            cache[key] = st
        return st

    def _eval_half(self, u_hat: torch.Tensor, f_mesh: FourierMesh, n_channel: int):
        """Right-hand side L u + N(u) of this operator on a rot-half state (B, C, modes) -> (out_hat, output channels).
        The body of ``__call__`` (operator/_base.py:753-790) between the transforms; also how a force operator handed to
        the pressure diagnostics is evaluated (dedicated/_navier_stokes.py:128-131, 186-189)."""
        if self._state_dict["f_mesh"] is not f_mesh or self._state_dict["n_channel"] != n_channel or self._lowered is None:
            self.register_mesh(f_mesh, n_channel)
        lo, B = self._lowered, u_hat.shape[0]
        if "map" in lo:
            return self._tf(B, n_channel).spectral_map(u_hat, lo["c_out"], lo["map"]), lo["c_out"]
        if "composite" in lo:
            return self._eval_composite(lo["composite"], u_hat, f_mesh, n_channel), 1
        if self.is_linear and not any(isinstance(t.coef, torch.Tensor) or t.kind == "linear_tensor" for t in self._terms_now):
            # a purely linear operator is a point-wise map too; this route also takes odd-order derivatives on 2-D/3-D
            # grids, whose complex symbol the time-stepping tables refuse
            if "linear_map" not in lo:
                lo["linear_map"] = self._lower_map(f_mesh, n_channel)
            c_out, terms = lo["linear_map"]
            return self._tf(B, n_channel).spectral_map(u_hat, c_out, terms), c_out
        st = getattr(self, "_rhs_stepper", None)
        if st is None or st.B != B or st.f_mesh is not f_mesh:
            st = self._build_integrator(1.0, B, rhs_only=True)
        return st.rhs_half(u_hat), n_channel

    def _eval_composite(self, t: _Term, u_hat: torch.Tensor, f_mesh: FourierMesh, n_channel: int) -> torch.Tensor:
        """Velocity2Pressure / Vorticity2Pressure: p = -lap^-1 div((u . grad) u - f)
        (dedicated/_navier_stokes.py:107-163, 166-217): [map to velocity ->] one evaluation of the convection program
        (which dealiases its input on load, :136-137/:193) -> pressure solve as a point-wise map."""
        B, d = u_hat.shape[0], f_mesh.n_dim
        force = t.params.get("external_force")
        conv_op = self.__dict__.get("_conv_op")
        if conv_op is None:
            conv_op = self._conv_op = Operator([_Term("convection")])
        conv_op._de_aliasing_rate = self._de_aliasing_rate
        conv_op._slab = self._slab
        vel = u_hat
        if t.kind == "vorticity2pressure":
            vel = self._tf(B, 1).spectral_map(u_hat, 2, [(0, 0, (0, 1, 0), 1, -1.0), (1, 0, (1, 0, 0), 1, 1.0)])
        conv, _ = conv_op._eval_half(vel, f_mesh, d)
        if force is not None:
            f_hat, _ = force._eval_half(u_hat, f_mesh, n_channel)      # the force sees the un-dealiased state
            conv = conv - f_hat
        e = [tuple(1 if i == a else 0 for i in range(3)) for a in range(d)]
        return self._tf(B, d).spectral_map(conv, 1, [(0, a, e[a], 1, -float(t.coef)) for a in range(d)])

    @staticmethod
    def _linear_core(t: _Term, f_mesh: FourierMesh, n_channel: int) -> torch.Tensor:
        if t.kind == "laplacian":                                                 # generic/_laplacian.py:12-15
            return torch.cat([f_mesh.laplacian()] * n_channel, dim=1)
        if t.kind == "biharmonic":                                                # generic/_biharmonic.py:13-16
            return torch.cat([f_mesh.laplacian() * f_mesh.laplacian()] * n_channel, dim=1)
        if t.kind == "spatial_derivative":                                        # generic/_spatial_derivative.py
            if n_channel != 1:
                raise ValueError("The SpatialDerivative operator only supports scalar field.")
            return f_mesh.grad(t.params["dim_index"], t.params["order"])
        if t.kind == "implicit_unit_source":                                      # generic/_source.py:14-17
            return torch.ones_like(f_mesh.bf(0))
        if t.kind == "linear_tensor":
            return t.params["L"].to(f_mesh.device)
        raise ValueError(t.kind)

    def _pre_check(self, u, u_fft, mesh):
        """operator/_base.py:528-579"""
        if u_fft is None and u is None:
            raise ValueError("Either u or u_fft should be given")
        if u_fft is not None and u is not None:
            assert u.shape == u_fft.shape, "The shape of u and u_fft should be the same"
        assert mesh is not None, "Mesh should be given"
        value = u if u is not None else u_fft
        if value.requires_grad and torch.is_grad_enabled() and not getattr(self, "_autograd_ok", False):
            raise NotImplementedError("gradients are available through integrate(u_0) and operator(u) with real-space "
                                      "inputs only (no u_fft / return_in_fourier / recorder); detach the input first")
        if not isinstance(mesh, FourierMesh):
            if isinstance(mesh, MeshGrid):
                mesh = FourierMesh(mesh, device=value.device, dtype=mesh.dtype)
            else:
                mesh = FourierMesh(mesh, device=value.device, dtype=value.dtype)
        n_channel = value.shape[1]
        assert len(value.shape) == mesh.n_dim + 2, \
            f"the value shape {tuple(value.shape)} is not compatible with mesh dim {mesh.n_dim}"
        for i in range(mesh.n_dim):
            expect = mesh.mesh_info[i][2]
            if i == 0 and getattr(self, "_slab", None) is not None:
                expect //= self._slab[1]                     # local x-slab of a slab-decomposed grid
            assert value.shape[i + 2] == expect, \
                f"Expect to have {expect} points in dim {i} but got {value.shape[i + 2]}"
        assert _same_device(value.device, mesh.device), \
            "The device of mesh {} and the device of value {} are not the same".format(mesh.device, value.device)
        assert self._value_mesh_check_func(len(value.shape) - 2, mesh.n_dim), \
            "Value and mesh do not match the requirement"
        return mesh, n_channel

    def _external_nonlinear(self, terms, n_channel: int):
        """sum_t coef_t * core_t(u_hat) for the host-composed cores (operator/_base.py:375-403): dealiased input for the
        cores that ask for it, one inverse transform shared by all of them, every transform on the library's passes."""
        d = self._state_dict["f_mesh"].n_dim
        def wants_dealiased(t):
            """True: the core reads the dealiased state; False: the state as it is; None: it masks on its own"""
            if t.kind == "custom_nonlinear":
                return bool(getattr(t.params["func"], "_dealiasing_swtich", True))
            if t.kind == "implicit_func_source":
                return bool(t.params.get("non_linear", True))
            return True if t.kind == "conservative_convection" else None
        need_masked = any(wants_dealiased(t) is True for t in terms)
        need_plain = any(wants_dealiased(t) is False for t in terms)
        f_mesh = self._state_dict["f_mesh"]
        pairs = [(a, c) for a in range(n_channel) for c in range(a, n_channel)]

        def scaled(r, coef, extra=1.0):
            """coef * r for a number or a per-sample tensor (B, 1, 1, ..) -> (B, 1, 1) on the half-spectrum layout"""
            if isinstance(coef, torch.Tensor):
                return r * (coef.reshape(-1, 1, 1).to(device=r.device, dtype=r.real.dtype) * extra)
            c = float(coef) * extra
            return r if c == 1.0 else r * c

        def evaluate(st, x_hat):
            x_d = st.mask_state(x_hat.clone()) if need_masked else None
            u_d = st.c2r(x_d) if need_masked else None
            u = st.c2r(x_hat) if need_plain else None
            out = None
            for t in terms:
                if t.kind == "implicit_func_source":                     # generic/_source.py:20-42
                    v = t.params["source_func"](u_d if t.params.get("non_linear", True) else u)
                    if v.shape != (st.B, st.C) + tuple(st.local_shape):
                        raise ValueError("ImplicitSource: source_func must keep the shape of its argument")
                    r = scaled(st.r2c(v), t.coef)
                elif t.kind == "custom_nonlinear":                       # operator/_base.py:375-403 with a user core
                    deal = wants_dealiased(t)
                    full = t.params["func"](st.half_to_full(x_d if deal else x_hat), getattr(st, "mesh", f_mesh),
                                            u_d if deal else u)
                    if tuple(full.shape) != (st.B, st.C) + tuple(st.local_shape):
                        raise ValueError("NonlinearFunc: the returned spectrum must keep the shape of its argument")
                    r = scaled(st.full_to_half(full), t.coef)
                elif t.kind == "ks_convection":                          # dedicated/_ks_convection.py:18-38 on a 1-D grid
                    g_hat = st.spectral_map(x_hat, 1, [(0, 0, (1, 0, 0), 0, 1.0)], dealias=True)      # i k phi_hat, dealiased
                    r = st.r2c(st.sym_outer(st.c2r(g_hat)))              # (phi_x)^2
                    if t.params.get("remove_mean", True):                # the mean spans batch and space: zero modes only
                        if r.requires_grad:
                            dc = torch.zeros_like(r)
                            dc[:, :, 0] = 1.0
                            r = r - dc * r[:, :, 0].mean()
                        else:
                            r[:, :, 0] -= r[:, :, 0].mean()
                    r = scaled(r, t.coef, 0.5)
                else:                                                    # generic/_conservative_convection.py:18-27
                    # gradient mode hands in a differentiable view of the same passes (autograd._HostView)
                    tf = st.other(len(pairs)) if hasattr(st, "other") else self._tf(st.B, len(pairs))
                    uu_hat = tf.r2c(st.sym_outer(u_d))
                    e = [tuple(1 if i == a else 0 for i in range(3)) for a in range(d)]
                    m = [(c, pairs.index((min(i, c), max(i, c))), e[i], 0, 1.0)
                         for c in range(n_channel) for i in range(d)]
                    r = scaled(tf.spectral_map(uu_hat, n_channel, m), t.coef)
                out = r if out is None else (out + r if r.requires_grad else out.add_(r))
            return out
        return evaluate

    def _build_integrator(self, dt: float, batch: int, tables: Optional[dict] = None, rhs_only: bool = False):
        """operator/_base.py:441-526: installs a FusedStepper in ``_state_dict['integrator']``.

        ``rhs_only`` plans the bare right-hand side (no ETD tables needed) and does not install it."""
        sd, lo = self._state_dict, self._lowered
        if rhs_only:
            name, cfg = "RK4", {}
        else:
            # an explicit source is a nonlinear core in the reference (operator/_base.py:994-1015): "auto" -> SETDRK4
            name, cfg = integrator_name(self._integrator, lo["program"] == _cabi.PROG_LINEAR and lo["source_hat"] is None
                                        and not lo["external"]), self._integrator_config
        if sd["f_mesh"].n_dim > 1:
            if has_imag(sd["linear_coef"]) or (tables is not None and any(has_imag(t) for t in tables.values())):
                # odd-order linear terms on a 2-D/3-D grid: the state leaves the Hermitian subspace on the Nyquist planes;
                # stepped as a pair of half spectra on the library's transform kernels (unrolled.py)
                if cfg.get("adaptive"):
                    raise NotImplementedError("adaptive Runge-Kutta stepping is host-synchronous and not part of the fused CUDA path")
                assert name != "ETDRK0" or rhs_only or (lo["program"] == _cabi.PROG_LINEAR and lo["source_hat"] is None
                                                        and not lo["external"]), "The ETDRK0 integrator only supports linear term"
                st = PairedSpectrumStepper(self, batch, name, dt, cfg, tables=tables)
                if rhs_only:
                    self._rhs_stepper = st
                else:
                    sd["integrator"] = st
                return st
        generic_rk = None
        if name in RK_TABLEAUS:               # explicit RK other than RK4: right-hand sides of an RK4-type plan + fsm_lincomb
            if cfg.get("adaptive"):
                raise NotImplementedError("adaptive Runge-Kutta stepping is host-synchronous and not part of the fused CUDA path")
            if lo.get("dyn_force") is not None or lo.get("force_hat") is not None:
                raise NotImplementedError("NSPressureConvection with an external force is not supported with the RK integrators")
            if self._slab is not None:
                raise NotImplementedError("only RK4 of the Runge-Kutta family runs on slab-decomposed grids")
            generic_rk, name, cfg = name, "RK4", {}
        if name == "ETDRK0":
            assert lo["program"] == _cabi.PROG_LINEAR and ((lo["source_hat"] is None and not lo["external"]) or rhs_only), \
                "The ETDRK0 integrator only supports linear term"
        try:
            extra = {}
            cls = FusedStepper
            if lo["external"]:
                cls, extra = HostComposedStepper, {"nonlinear": self._external_nonlinear(lo["external"], sd["n_channel"])}
            elif lo["dyn_force"] is not None:
                if name == "RK4" and not rhs_only:
                    raise NotImplementedError("NSPressureConvection with an external force is not supported with the RK integrators")
                if self._slab is not None:
                    raise NotImplementedError("state-dependent external forces are not available on slab-decomposed grids")
                if sd["f_mesh"].dtype == torch.float64:
                    # the force acts on the un-dealiased state, Nyquist planes included, where the reference's full C2C
                    # state carries anti-Hermitian content that the pressure symbol turns into real-space output
                    # (SURVEY.md H1); a half-spectrum state cannot hold it. Measured against the reference: 1e-8 relative
                    # after three steps -- below fp32 rounding, four orders above the fp64 tolerance.
                    raise NotImplementedError("NSPressureConvection with a state-dependent external force reproduces the "
                                              "reference to ~1e-8 only (Nyquist-plane content of its full spectrum): "
                                              "available in fp32, refused in fp64")
                cls, extra = DynamicForceStepper, {"force": lo["dyn_force"]}
            st = cls(sd["f_mesh"], batch, sd["n_channel"], lo["program"], name, dt, sd["linear_coef"],
                     lo["nl_coef"], lo["source_hat"], lo["kmax"], lo["ks_remove_mean"], cfg,
                     chunk=self._chunk, tables=tables, slab=self._slab, lanes=self._lanes,
                     force_hat=lo["force_hat"], nl_coef_b=lo.get("nl_coef_b"), **extra)
        except torch.cuda.OutOfMemoryError as e:
            raise RuntimeError(os.linesep.join([
                "Cuda out of memory when building the integrator.",
                "Original error message: {}".format(str(e)),
                "Please try to use a smaller mesh or a low-order integrator."]))
        st.ks_group = getattr(self, "_ensemble_group", None)
        if generic_rk is not None:
            st = ExplicitRKStepper(st, generic_rk, dt)
        if rhs_only:
            self._rhs_stepper = st
        else:
            sd["integrator"] = st
        return st

    def _stepper_for(self, value, mesh, dt):
        if self._state_dict["f_mesh"] is None or mesh is not None or self._lowered is None:
            mesh, n_channel = self._pre_check(value[0], value[1], mesh if mesh is not None else self._state_dict["f_mesh"])
            self.register_mesh(mesh, n_channel)
        else:
            self._pre_check(value[0], value[1], self._state_dict["f_mesh"])
        if "map" in self._lowered or "composite" in self._lowered:
            raise NotImplementedError("Grad/Div/Curl and the velocity/pressure diagnostics change the channel count: they "
                                      "can be evaluated (operator(u)) but not integrated in time")
        v = value[0] if value[0] is not None else value[1]
        st = self._state_dict["integrator"]
        if st is None or st.dt != dt or st.B != v.shape[0]:
            st = self._build_integrator(dt, v.shape[0])
        return st

    # ---- the hot path ----------------------------------------------------------------------------------
    def integrate(self, u_0: Optional[torch.Tensor] = None, u_0_fft: Optional[torch.Tensor] = None, dt: float = 1,
                  step: int = 1, mesh=None, progressive: bool = False, trajectory_recorder=None,
                  return_in_fourier: bool = False):
        """operator/_base.py:676-751 — same signature and return conventions."""
        if u_0 is not None and self._needs_graph(u_0) and u_0_fft is None and trajectory_recorder is None \
                and not return_in_fourier and not progressive:
            return self._integrate_with_grad(u_0, dt, step, mesh)
        st = self._stepper_for((u_0, u_0_fft), mesh, dt)
        try:
            u_hat = st.r2c(u_0) if u_0_fft is None else st.full_to_half(u_0_fft)
            bar = None
            if progressive:
                from tqdm.auto import tqdm
                bar = tqdm(total=step, desc="Integrating")
            if trajectory_recorder is None and bar is None:
                st.step_half(u_hat, step)
            else:
                i = 0
                if hasattr(trajectory_recorder, "prepare"):          # per-sample frame draws need the batch size
                    trajectory_recorder.prepare(st.B)
                control = getattr(trajectory_recorder, "control_func", None) if trajectory_recorder is not None else None
                # recorders of this package take physical frames straight from the C2R pass; anything else gets
                # the reference's full-spectrum frame (traj_recorder.py:46-55)
                real_frames = (trajectory_recorder is not None and not return_in_fourier
                               and getattr(trajectory_recorder, "accepts_real_frames", False))

                def emit(k):
                    if real_frames:
                        trajectory_recorder.record_real(k, st.c2r(u_hat))
                    else:
                        trajectory_recorder.record(k, st.half_to_full(u_hat))
                while i < step:
                    if trajectory_recorder is not None and (control is None or control(i)):
                        emit(i)
                    # fuse the steps up to the next frame the recorder may want
                    j = i + 1
                    if control is not None or trajectory_recorder is None:
                        limit = min(step, i + (max(1, step // 100) if bar is not None else step))
                        while j < limit and not (control is not None and control(j)):
                            j += 1
                    st.step_half(u_hat, j - i)
                    if bar is not None:
                        bar.update(j - i)
                    i = j
            if bar is not None:
                bar.close()
            if trajectory_recorder is not None:
                if control is None or control(step):
                    emit(step)
                trajectory_recorder.return_in_fourier = return_in_fourier
                return trajectory_recorder.trajectory
            return st.half_to_full(u_hat) if return_in_fourier else st.c2r(u_hat)
        except torch.cuda.OutOfMemoryError as e:
            raise RuntimeError(os.linesep.join([
                "Cuda out of memory when integrating the operator.",
                "Original error message: {}".format(str(e)),
                "Please try to use a smaller mesh or a low-order integrator."]))

    def integrate_stream(self, batches, dt: float = 1, step: int = 1, mesh=None, out=None):
        """Ensemble streaming (dataset generation): every element of ``batches`` is a HOST tensor ``(B, C, *grid)``
        (pinned memory for asynchronous copies) that is uploaded, advanced by ``step`` steps exactly as
        ``integrate(u_0, dt=dt, step=step)`` would, and downloaded into the matching element of ``out`` (host
        tensors; allocated pinned when omitted). Upload, device work and download of consecutive batches run on
        three CUDA streams with double-buffered device staging, so the PCIe copies (both directions at once)
        hide behind each other and behind the step. Returns the list of host results; the calling stream is
        ordered after the last download."""
        batches = list(batches)
        if not batches:
            return []
        first = batches[0]
        if first.device.type != "cpu":
            raise ValueError("integrate_stream takes host tensors; use integrate() for device-resident states")
        f_mesh = mesh if mesh is not None else self._state_dict["f_mesh"]
        if f_mesh is None:
            raise ValueError("Mesh should be given")
        if not isinstance(f_mesh, FourierMesh):
            f_mesh = FourierMesh(f_mesh) if isinstance(f_mesh, MeshGrid) else FourierMesh(f_mesh, dtype=first.dtype)
        dev = f_mesh.device
        if dev.type != "cuda":
            raise ValueError("integrate_stream needs a mesh on a CUDA device")
        stage_in = [torch.empty(first.shape, dtype=first.dtype, device=dev) for _ in range(2)]
        st = self._stepper_for((stage_in[0], None), mesh, dt)
        stage_out = [torch.empty((st.B, st.C) + tuple(st.local_shape), dtype=st.rdtype, device=dev) for _ in range(2)]
        if out is None:
            out = [torch.empty(stage_out[0].shape, dtype=st.rdtype, pin_memory=True) for _ in batches]
        out = list(out)
        if len(out) != len(batches):
            raise ValueError("need one output tensor per batch")
        cur = torch.cuda.current_stream(dev)
        if getattr(self, "_io_streams", None) is None or self._io_streams[0].device != dev:
            self._io_streams = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
        up, down = self._io_streams
        start = torch.cuda.Event()
        start.record(cur)
        up.wait_event(start)
        down.wait_event(start)
        in_free = [None, None]      # staging input consumed by r2c
        out_free = [None, None]     # staging output downloaded
        for i, host in enumerate(batches):
            s = i & 1
            if tuple(host.shape) != tuple(first.shape):
                raise ValueError("all batches must have the same shape")
            with torch.cuda.stream(up):
                if in_free[s] is not None:
                    up.wait_event(in_free[s])
                stage_in[s].copy_(host, non_blocking=True)
                ev_up = torch.cuda.Event()
                ev_up.record(up)
            cur.wait_event(ev_up)
            u_hat = st.r2c(stage_in[s])
            in_free[s] = torch.cuda.Event()
            in_free[s].record(cur)
            st.step_half(u_hat, step)
            if out_free[s] is not None:
                cur.wait_event(out_free[s])
            st.c2r(u_hat, out=stage_out[s])
            ev_done = torch.cuda.Event()
            ev_done.record(cur)
            with torch.cuda.stream(down):
                down.wait_event(ev_done)
                out[i].copy_(stage_out[s], non_blocking=True)
                out_free[s] = torch.cuda.Event()
                out_free[s].record(down)
        for ev in out_free:
            if ev is not None:
                cur.wait_event(ev)
        # the staging buffers are referenced by work queued on the side streams
        for t in stage_in + stage_out:
            t.record_stream(up)
            t.record_stream(down)
        return out

    def __call__(self, u: Optional[torch.Tensor] = None, u_fft: Optional[torch.Tensor] = None, mesh=None,
                 return_in_fourier: bool = False):
        """operator/_base.py:753-790 — evaluate L u + N(u) once."""
        if u is not None and self._needs_graph(u) and u_fft is None and not return_in_fourier:
            return self._call_with_grad(u, mesh)
        if self._state_dict["f_mesh"] is None or mesh is not None or self._lowered is None:
            m, n_channel = self._pre_check(u, u_fft, mesh if mesh is not None else self._state_dict["f_mesh"])
            self.register_mesh(m, n_channel)
        else:
            self._pre_check(u, u_fft, self._state_dict["f_mesh"])
        f_mesh, n_channel = self._state_dict["f_mesh"], self._state_dict["n_channel"]
        v = u if u is not None else u_fft
        B = v.shape[0]
        lo = self._lowered
        if "map" in lo or "composite" in lo:
            st_in, st_out = self._tf(B, n_channel), self._tf(B, lo["c_out"])
        elif self.is_linear and not any(isinstance(t.coef, torch.Tensor) or t.kind == "linear_tensor" for t in self._terms_now):
            st_in = st_out = self._tf(B, n_channel)
        else:
            st_in = getattr(self, "_rhs_stepper", None)
            if st_in is None or st_in.B != B or st_in.f_mesh is not f_mesh:
                st_in = self._build_integrator(1.0, B, rhs_only=True)
            st_out = st_in
        u_hat = st_in.r2c(u) if u_fft is None else st_in.full_to_half(u_fft)
        out, _ = self._eval_half(u_hat, f_mesh, n_channel)
        return st_out.half_to_full(out) if return_in_fourier else st_out.c2r(out)


    # ---- differentiable paths -------------------------------------------------------------------
    def _params_require_grad(self) -> bool:
        """Tensor-valued coefficients or explicit sources that are being optimised (inverse problems)."""
        for t in self.terms:
            if isinstance(t.coef, torch.Tensor) and t.coef.requires_grad:
                return True
            src = t.params.get("source")
            if isinstance(src, torch.Tensor) and src.requires_grad:
                return True
            core = t.params.get("generator")          # a user-defined core that is an nn.Module (learned closure)
            if isinstance(core, torch.nn.Module) and any(p.requires_grad for p in core.parameters()):
                return True
        return False

    def _needs_graph(self, u) -> bool:
        return torch.is_grad_enabled() and (u.requires_grad or self._params_require_grad())

    def _fresh_graph(self):
        """Parameters that require grad: symbol, tables and sources are rebuilt for every call, so that each result
        carries its own graph to them (a cached plan would hold the graph of an earlier call)."""
        if self._params_require_grad():
            self._lowered = None
            self._state_dict["integrator"] = None
            self._rhs_stepper = None
            return True
        return False

    def _call_with_grad(self, u, mesh):
        """``operator(u)`` with ``u.requires_grad``: linear operators and point-wise spectral maps (Grad, Div, Curl,
        Vorticity2Velocity, sums with linear cores) are differentiated by running the adjoint map on the cotangent;
        nonlinear operators through gradient mode (autograd.py)."""
        params = self._fresh_graph()
        self._autograd_ok = True
        try:
            if self._state_dict["f_mesh"] is None or mesh is not None or self._lowered is None:
                m, n_channel = self._pre_check(u, None, mesh if mesh is not None else self._state_dict["f_mesh"])
                self.register_mesh(m, n_channel)
            else:
                self._pre_check(u, None, self._state_dict["f_mesh"])
        finally:
            self._autograd_ok = False
        lo, n_channel, B = self._lowered, self._state_dict["n_channel"], u.shape[0]
        if "map" in lo:
            if params:
                raise NotImplementedError("gradients with respect to the coefficients of Grad/Div/Curl-type operators are "
                                          "not available on the CUDA path")
            c_out, terms = lo["c_out"], lo["map"]
        elif "composite" not in lo and self.is_linear and not any(isinstance(t.coef, torch.Tensor) for t in self._terms_now):
            c_out, terms = self._lower_map(self._state_dict["f_mesh"], n_channel)
        elif "composite" in lo:
            raise NotImplementedError("Velocity2Pressure / Vorticity2Pressure are not differentiable on the CUDA path; "
                                      "detach their input first")
        else:                       # nonlinear (or per-sample linear) right-hand side: unrolled on the library's passes
            st = getattr(self, "_rhs_stepper", None)
            if st is None or st.B != B or st.f_mesh is not self._state_dict["f_mesh"]:
                st = self._build_integrator(1.0, B, rhs_only=True)
            if hasattr(st, "evaluate_with_grad"):
                return st.evaluate_with_grad(u)
            return GradientMode(self, st).evaluate(u)
        adj = _adjoint_terms(terms)

        def run(x, c_in, c_to, tt):
            st_in, st_to = self._tf(B, c_in), self._tf(B, c_to)
            return st_to.c2r(st_in.spectral_map(st_in.r2c(x), c_to, tt))
        return _LinearFn.apply(u, lambda x: run(x, n_channel, c_out, terms), lambda g: run(g, c_out, n_channel, adj))

    def _integrate_with_grad(self, u_0, dt, step, mesh):
        """``integrate(u_0)`` with ``u_0.requires_grad``. A purely linear operator with real tables (exp(L dt) real:
        even-order terms) is self-adjoint, so the cotangent is integrated by the same fused plan; everything else goes
        through gradient mode (autograd.py)."""
        params = self._fresh_graph()
        self._autograd_ok = True
        try:
            st = self._stepper_for((u_0, None), mesh, dt)
        finally:
            self._autograd_ok = False
        lo = self._lowered
        if params:
            self._state_dict["integrator"] = None       # the plan's tables carry this call's graph: not reused
        if hasattr(st, "integrate_with_grad"):          # complex symbol on a 2-D/3-D grid (unrolled.py)
            return st.integrate_with_grad(u_0, step)
        if lo.get("program") != _cabi.PROG_LINEAR or lo.get("source_hat") is not None or lo.get("external") or st.complex_tables \
                or st.integrator != "ETDRK0" or params:
            # nonlinear operators: the step is unrolled on the library's transform and symbol kernels, each with its
            # adjoint pass as backward (torchfsm_b200/autograd.py)
            return GradientMode(self, st).integrate(u_0, step)

        def run(x):
            return st.c2r(st.step_half(st.r2c(x), step))
        return _LinearFn.apply(u_0, run, run)


class Operator(OperatorLike):
    pass


def _solve(self, b: Optional[torch.Tensor] = None, b_fft: Optional[torch.Tensor] = None, mesh=None,
           n_channel: Optional[int] = None, return_in_fourier: bool = False):
    """Solve the linear equation ``A x = b`` (mirror of ``_InverseSolveMixin.solve``, operator/_base.py:217-262):
    ``x_hat = b_hat * where(L == 0, 1, 1/L)``. Runs on the CUDA library: r2c, one table multiply, c2r."""
    value = b if b is not None else b_fft
    if value is None:
        raise ValueError("Either b or b_fft should be given")
    if mesh is None:
        assert self._state_dict["f_mesh"] is not None, "Mesh and n_channel should be given when calling solve"
        mesh = self._state_dict["f_mesh"]
    f_mesh, c = self._pre_check(b, b_fft, mesh)
    n_channel = c if n_channel is None else n_channel
    self.register_mesh(f_mesh, n_channel)            # user-defined generators are resolved here
    if not self.is_linear:
        raise NotImplementedError("solve is only defined for linear operators")
    L = self._state_dict["linear_coef"]
    inv = torch.where(L == 0, 1.0, 1 / L)                                   # operator/_base.py:250-255
    if f_mesh.n_dim > 1 and has_imag(inv):        # odd-order terms on a 2-D/3-D grid: 1/L(k) and 1/L(-k) are no conjugates
        st = PairedSpectrumStepper(self, value.shape[0], "ETDRK0", 1.0, {}, tables={"exp": inv})
    else:
        st = FusedStepper(f_mesh, value.shape[0], n_channel, _cabi.PROG_LINEAR, "ETDRK0", 1.0, None, 0.0, None,
                          [n // 2 for n in f_mesh.shape], True, {}, tables={"exp": inv})
    x_hat = st.r2c(b) if b_fft is None else st.full_to_half(b_fft)
    st.step_half(x_hat, 1)
    return st.half_to_full(x_hat) if return_in_fourier else st.c2r(x_hat)


OperatorLike.solve = _solve


class LinearOperator(OperatorLike):
    pass


class NonlinearOperator(OperatorLike):
    pass


# ---- generator wrappers with the reference's class names ------------------------------------------
def Laplacian() -> Operator:
    """operator/generic/_laplacian.py:17-25"""
    return Operator([_Term("laplacian")])


def Biharmonic() -> Operator:
    """operator/generic/_biharmonic.py:19-28"""
    return Operator([_Term("biharmonic")])


def SpatialDerivative(dim_index: int, order: int) -> Operator:
    """operator/generic/_spatial_derivative.py:42-55"""
    return Operator([_Term("spatial_derivative", 1, {"dim_index": dim_index, "order": order})])


def ImplicitSource(source_func=None, non_linear: bool = True) -> Operator:
    """operator/generic/_source.py:45-74. The unit form is a linear core; ``source_func`` (a callable on the physical
    field) becomes a host-composed nonlinear term: c2r -> source_func -> r2c on the library's passes."""
    if source_func is not None:
        return Operator([_Term("implicit_func_source", 1, {"source_func": source_func, "non_linear": non_linear})])
    return Operator([_Term("implicit_unit_source")])


def ExplicitSource(source: torch.Tensor) -> Operator:
    """operator/_base.py:1018-1028"""
    return Operator([_Term("explicit_source", 1, {"source": source})])


def Convection() -> Operator:
    """operator/generic/_convection.py:66-76"""
    return Operator([_Term("convection")])


def KSConvection(remove_mean: bool = True) -> Operator:
    """operator/dedicated/_ks_convection.py:54-67"""
    return Operator([_Term("ks_convection", 1, {"remove_mean": remove_mean})])


def VorticityConvection() -> Operator:
    """operator/dedicated/_navier_stokes.py:61-72"""
    return Operator([_Term("vorticity_convection")])


def NSPressureConvection(external_force=None) -> Operator:
    """operator/dedicated/_navier_stokes.py:257-268"""
    return Operator([_Term("ns_pressure_convection", 1, {"external_force": external_force})])


def ConservativeConvection() -> Operator:
    """operator/generic/_conservative_convection.py:46-55: div(u u); host-composed from c2r, the symmetric products,
    r2c of the d(d+1)/2 distinct products and the divergence as a spectral map."""
    return Operator([_Term("conservative_convection")])


def Grad() -> Operator:
    """operator/generic/_grad.py:30-38"""
    return Operator([_Term("grad")])


def Div() -> Operator:
    """operator/generic/_div.py:42-53"""
    return Operator([_Term("div")])


def Curl() -> Operator:
    """operator/generic/_curl.py:78-93"""
    return Operator([_Term("curl")])


def Vorticity2Velocity() -> Operator:
    """operator/dedicated/_navier_stokes.py:107-116"""
    return Operator([_Term("vorticity2velocity")])


def Vorticity2Pressure(external_force: Optional[OperatorLike] = None) -> Operator:
    """operator/dedicated/_navier_stokes.py:166-176"""
    return Operator([_Term("vorticity2pressure", 1, {"external_force": external_force})])


def Velocity2Pressure(external_force: Optional[OperatorLike] = None) -> Operator:
    """operator/dedicated/_navier_stokes.py:220-229"""
    return Operator([_Term("velocity2pressure", 1, {"external_force": external_force})])


def run_operators(u: torch.Tensor, operators: Sequence[OperatorLike], mesh):
    """operator/__init__.py:12-31 — apply several operators to one field; the forward transform is shared."""
    if not operators:
        return iter(())
    first = operators[0]
    f_mesh, n_channel = first._pre_check(u, None, mesh)
    first.register_mesh(f_mesh, n_channel)
    st = first._tf(u.shape[0], n_channel)
    u_hat = st.r2c(u)

    def run(op):
        out, c_out = op._eval_half(u_hat, f_mesh, n_channel)
        return op._tf(u.shape[0], c_out).c2r(out)
    return map(run, operators)
