"""Row-partitioned FP32 SpMM of one large mesh across the GPUs of a node.

Rank r owns the contiguous slab of node rows [bounds[r], bounds[r+1]) of the block CSR (the node
numbering is lexicographic in the coordinates, so a slab is a geometric slab and couples only to its
two neighbours).  Every rank keeps its slab of each dense block in peer-visible memory (cudaMalloc +
CUDA IPC); the SpMM kernel (csrc/precond32.cu, PEER variant of k_spmm32v) reads the halo rows straight
from the neighbours' memory over NVLink -- there is no all-gather and no staging copy.  The only
cross-rank ordering needed is a barrier between the step that writes a block and the step that gathers
it, which the caller issues on the stream (torch.distributed).

Reference behaviour replaced: nothing one-to-one -- the reference is single-GPU;
a torch port of `stiff_matrix @ U` across GPUs would all-gather U (n x k) every product.
"""
import ctypes as C
from typing import List

import torch
import torch.distributed as dist

from .. import _lib, native


def slab_bounds(n_nodes: int, world: int) -> List[int]:
    """Contiguous, near-equal slabs of node rows: bounds[r] .. bounds[r+1]."""
    if world < 1 or world > 8:
        raise ValueError("row partition supports 1..8 ranks (one NVLink node)")
    if n_nodes < world:
        raise ValueError(f"{n_nodes} node rows cannot be split over {world} ranks")
    base, rem = divmod(n_nodes, world)
    out = [0]
    for r in range(world):
        out.append(out[-1] + base + (1 if r < rem else 0))
    return out


def owner_of(nodes: torch.Tensor, bounds: List[int]) -> torch.Tensor:
    """Rank that owns each node id."""
    b = torch.as_tensor(bounds[1:-1], dtype=nodes.dtype, device=nodes.device)
    return torch.searchsorted(b, nodes, right=True)


def packed_column_map(n_nodes: int, bounds: List[int], device) -> torch.Tensor:
    """colmap[j] = owner(j) << 28 | (j - bounds[owner(j)]) as int32 bit pattern (uint32 in the kernel)."""
    if max(b1 - b0 for b0, b1 in zip(bounds[:-1], bounds[1:])) >= 1 << 28:
        raise ValueError("slab too large for the 28-bit local index")
    j = torch.arange(n_nodes, dtype=torch.int64, device=device)
    own = owner_of(j, bounds)
    start = torch.as_tensor(bounds, dtype=torch.int64, device=device)[own]
    packed = (own << 28) | (j - start)
    return packed.to(torch.int32)          # world <= 8 keeps bit 31 clear


class _PeerArray:
    """A float32 (rows, cols) array in peer-visible memory, exposed to torch without a copy."""

    def __init__(self, ptr, rows, cols):
        self.ptr, self.rows, self.cols = ptr, rows, cols
        self.__cuda_array_interface__ = {"shape": (rows, cols), "typestr": "<f4", "data": (ptr, False), "version": 3,
                                         "strides": None}


class RowPartition:
    """This rank's slab of K (FP32 records) plus `nbuf` peer-visible dense blocks of `ncols` columns."""

    def __init__(self, pattern, Kval, ncols, nbuf=2, Mblk=None, shift=0.0, group=None):
        lib = _lib.load()
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        dev = Kval.device
        self.device = dev
        self.ncols = int(ncols)
        self.bounds = slab_bounds(pattern.n_nodes, self.world)
        r0, r1 = self.bounds[self.rank], self.bounds[self.rank + 1]
        self.row0, self.n_local = r0, r1 - r0
        brow = pattern.brow
        b0, b1 = int(brow[r0]), int(brow[r1])
        self.nnzb_local = b1 - b0
        self.brow_local = (brow[r0:r1 + 1] - b0).contiguous()
        colmap = packed_column_map(pattern.n_nodes, self.bounds, dev)
        self.rec = torch.empty((lib.ds_k32_record_bytes(self.nnzb_local) + 15) // 16 * 4, dtype=torch.int32, device=dev)
        self.invD = torch.empty(self.n_local * 9, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.ds_k32_pack_slab(native._p(self.brow_local), native._p(pattern.bcol[b0:b1].contiguous()),
                                            self.n_local, self.nnzb_local, r0,
                                            native._p(Kval[9 * b0:9 * b1].contiguous()),
                                            native._p(Mblk[b0:b1].contiguous()) if Mblk is not None else None,
                                            float(shift), native._p(colmap), native._p(self.rec), native._p(self.invD),
                                            native._stream()), "ds_k32_pack_slab")
            # peer-visible dense blocks: one allocation, nbuf slabs of (3 * max_local, ncols) fp32
            self.max_local = max(b - a for a, b in zip(self.bounds[:-1], self.bounds[1:]))
            self.buf_elems = 3 * self.max_local * self.ncols
            nbytes = 4 * self.buf_elems * nbuf
            ptr = C.c_void_p()
            handle = (C.c_ubyte * 64)()
            _lib.check(lib.ds_peer_alloc(nbytes, C.byref(ptr), handle), "ds_peer_alloc")
            self._own_ptr = ptr.value
            handles = [None] * self.world
            if self.world > 1:
                dist.all_gather_object(handles, bytes(handle), group=group)
            else:
                handles[0] = bytes(handle)
            self._peer_ptrs = []
            for r, h in enumerate(handles):
                if r == self.rank:
                    self._peer_ptrs.append(self._own_ptr)
                else:
                    p = C.c_void_p()
                    hb = (C.c_ubyte * 64).from_buffer_copy(h)
                    _lib.check(lib.ds_peer_open(hb, C.byref(p)), "ds_peer_open")
                    self._peer_ptrs.append(p.value)
        self.nbuf = nbuf
        self.blocks = [torch.as_tensor(_PeerArray(self._own_ptr + 4 * self.buf_elems * k, 3 * self.n_local, self.ncols),
                                       device=dev) for k in range(nbuf)]

    def barrier(self):
        """Order the ranks on the current stream (NCCL all-reduce of one element; no host sync)."""
        if self.world > 1:
            if not hasattr(self, "_tok"):
                self._tok = torch.zeros(1, device=self.device)
            dist.all_reduce(self._tok, group=self.group)

    def spmm(self, src, out, mode=0, R=None, Zprev=None, ab=0.0, cc=0.0):
        """out (local tensor or one of self.blocks) = op(A_slab, gathered block `src` (index into
        self.blocks on every rank)).  mode as native.spmm32."""
        lib = _lib.load()
        parts = (C.c_void_p * self.world)(*[p + 4 * self.buf_elems * src for p in self._peer_ptrs])
        with torch.cuda.device(self.device):
            _lib.check(lib.ds_spmm32_rowpart(int(mode), native._p(self.brow_local), native._p(self.rec), self.n_local,
                                             self.ncols, parts, self.world, self.rank, native._p(R),
                                             native._p(self.invD) if mode == 2 else None, native._p(Zprev),
                                             native._p(out), float(ab), float(cc), native._stream()),
                       "ds_spmm32_rowpart")
        return out

    def close(self):
        lib = _lib.load()
        if getattr(self, "_peer_ptrs", None):
            torch.cuda.synchronize(self.device)
            if self.world > 1:
                dist.barrier(group=self.group)
            for r, p in enumerate(self._peer_ptrs):
                if r != self.rank:
                    lib.ds_peer_close(C.c_void_p(p))
            if self.world > 1:
                dist.barrier(group=self.group)
            self.blocks = []
            lib.ds_peer_free(C.c_void_p(self._own_ptr))
            self._peer_ptrs = None
