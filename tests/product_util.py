"""Build torchfsm_b200 operators from golden-fixture specs (shared by emulator and GPU tests)."""
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_LIB = os.path.join(ROOT, "tests", "emu", "_build", "libfsm_emu.so")


def build_emulator():
    import subprocess
    subprocess.run(["make", "-s", "-j8", "-C", os.path.join(ROOT, "torchfsm_b200", "csrc"), "emu"], check=True)
    return EMU_LIB


def product_operator(terms, sources=None):
    """terms = [(kind, coef, params)] -> torchfsm_b200 Operator built from its public classes."""
    import torchfsm_b200 as fsm
    op = None
    for i, (kind, coef, params) in enumerate(terms):
        if kind == "laplacian":
            t = fsm.Laplacian()
        elif kind == "biharmonic":
            t = fsm.Biharmonic()
        elif kind == "spatial_derivative":
            t = fsm.SpatialDerivative(params["dim_index"], params["order"])
        elif kind == "implicit_unit_source":
            t = fsm.ImplicitSource()
        elif kind == "convection":
            t = fsm.Convection()
        elif kind == "ks_convection":
            t = fsm.KSConvection(params.get("remove_mean", True))
        elif kind == "vorticity_convection":
            t = fsm.VorticityConvection()
        elif kind == "ns_pressure_convection":
            t = fsm.NSPressureConvection()
        elif kind == "explicit_source":
            t = fsm.ExplicitSource(params["source"])
        else:
            raise ValueError(kind)
        t = coef * t
        op = t if op is None else op + t
    return op


def integrator_enum(name):
    import torchfsm_b200 as fsm
    if name == "auto":
        return "auto"
    for enum in (fsm.ETDRKIntegrator, fsm.SETDRKIntegrator, fsm.RKIntegrator):
        if name in enum.__members__:
            return enum[name]
    raise ValueError(name)


def product_from_golden(g, device):
    """Operator + mesh + u0 tensor for a golden fixture on `device`."""
    import torchfsm_b200 as fsm
    spec = g["spec"]
    dtype = torch.float32 if spec["dtype"] == "float32" else torch.float64
    terms = []
    for i, (kind, coef, params) in enumerate(spec["terms"]):
        params = dict(params)
        if kind == "explicit_source":
            params["source"] = torch.from_numpy(g[f"src_{i}"]).to(device)
        terms.append((kind, coef, params))
    op = product_operator(terms)
    op.set_integrator(integrator_enum(spec["integrator"]))
    mesh = fsm.MeshGrid([tuple(m) for m in spec["mesh"]], device=device, dtype=dtype)
    u0 = torch.from_numpy(g["u0"]).to(device)
    return op, mesh, u0


class SharedFilePeers:
    """Peer provider for the CPU tests of the direct slab exchange (torchfsm_b200/peer.py protocol): every
    rank's receive buffer is a shared memory-mapped file that all (emulator) processes map, the barrier is
    the gloo barrier. On GPUs the same role is played by torch symmetric memory."""

    def __init__(self, rank, world, directory):
        self.rank, self.world, self.dir = rank, world, directory
        self._n = 0
        self._keep = []

    def alloc(self, numel, dtype, device):
        import torch
        import torch.distributed as dist
        nbytes = int(numel) * torch.empty((), dtype=dtype).element_size()
        idx, self._n = self._n, self._n + 1
        path = lambda r: os.path.join(self.dir, f"recv{idx}_rank{r}.bin")   # noqa: E731
        with open(path(self.rank), "wb") as f:
            f.truncate(nbytes)
        dist.barrier()
        maps = [torch.from_file(path(r), shared=True, size=nbytes, dtype=torch.uint8) for r in range(self.world)]
        self._keep.append([m.view(dtype) for m in maps])
        return maps[self.rank].view(dtype), [m.data_ptr() for m in maps], idx

    def remote(self, index, rank):
        return self._keep[index][rank]

    def barrier(self, index=0, channel=0):
        import torch.distributed as dist
        dist.barrier()
