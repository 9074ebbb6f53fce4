"""Peer-visible receive buffers for the direct exchange of slab-decomposed grids.

The forward-y and inverse-x kernels store straight into the receive buffers of the destination ranks
(`fsm_slab_peers` in include/fsm_b200.h), so an exchange is only a barrier across ranks on the stream.
A provider hands out buffers that every rank can address plus that barrier:

    alloc(numel, dtype, device) -> (local tensor, [address of rank r's buffer as seen from this process], index)
    remote(index, rank)         -> tensor view of rank `rank`'s buffer number `index` (allocation order)
    barrier(index=0, channel=0) -> all ranks' earlier work on their current streams is visible afterwards

Two exchange paths use them (`OperatorLike.set_slab_decomposition(exchange=...)`):
  "store": the kernels store into the peers' buffers themselves (no send buffer at all);
  "dma":   the kernels fill a local rank-blocked send buffer and each block is pushed into its owner's receive
           buffer by the copy engines (peer-to-peer cudaMemcpyAsync), which leaves every SM to the transforms.

`SymmetricMemoryPeers` is the CUDA provider: torch symmetric memory (NVLink peer mappings inside one NVSwitch
domain, device-side signal barrier). PyTorch is plumbing here; the data path is this repo's kernels.
"""


class SymmetricMemoryPeers:
    def __init__(self, group):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self._symm_mem = symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self._handles = []
        self._tensors = []
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")            # deprecated (a no-op) in recent releases, required in older ones
            try:
                symm_mem.enable_symm_mem_for_group(self.group.group_name)
            except Exception:
                pass

    def alloc(self, numel, dtype, device):
        t = self._symm_mem.empty(int(numel), dtype=dtype, device=device)
        hdl = self._symm_mem.rendezvous(t, self.group)
        self._handles.append(hdl)
        self._tensors.append(t)
        return t, [int(p) for p in hdl.buffer_ptrs], len(self._handles) - 1

    def remote(self, index, rank):
        hdl, t = self._handles[index], self._tensors[index]
        return hdl.get_buffer(rank, t.shape, t.dtype)

    def barrier(self, index=0, channel=0):
        self._handles[index].barrier(channel=channel)
