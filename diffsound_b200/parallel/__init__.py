"""Multi-GPU plumbing of the modal-analysis path: one process per GPU,
torch.distributed for rendezvous / barriers / result gathers.

* `sweep`:   independent candidates (thickness / material / morph sweeps of the reference's
             experiments) sharded across ranks, no data-path collective.
* `rowpart`: one large mesh, contiguous node-row slabs of the block CSR per rank; the SpMM reads the
             halo rows of the dense block straight from the peers' memory over NVLink.
"""
from .sweep import shard_indices, gather_ordered, sweep_modal_solves  # noqa: F401
from .rowpart import slab_bounds, owner_of, RowPartition  # noqa: F401
