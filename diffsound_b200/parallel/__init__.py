"""Multi-GPU plumbing of the modal-analysis path: one process per GPU,
torch.distributed for rendezvous / barriers / result gathers.

* `sweep`:   independent candidates (thickness / material / morph sweeps of the reference's
             experiments) sharded across ranks, no data-path collective.
* `synth`:   modal synthesis sharded over the batch axis; one all-reduce of the shared per-mode gradients.
* `rowpart`: one large mesh, contiguous node-row slabs of the block CSR per rank; the SpMM reads the
             halo rows of the dense block straight from the peers' memory over NVLink.
"""
from .sweep import shard_indices, gather_ordered, sweep_modal_solves  # noqa: F401
from .synth import batch_slice, sharded_modal_synth, allreduce_shared_grads  # noqa: F401
from .rowpart import slab_bounds, owner_of, RowPartition  # noqa: F401
