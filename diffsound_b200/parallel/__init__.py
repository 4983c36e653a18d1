"""Multi-GPU plumbing of the modal-analysis path: one process per GPU,
torch.distributed for rendezvous / barriers / result gathers.

* `sweep`:   independent candidates (thickness / material / morph sweeps of the reference's
             experiments) sharded across ranks, no data-path collective; round-robin, or on demand from a shared
             counter when the candidates differ widely in cost (`WorkQueue`).
* `synth`:   modal synthesis sharded over the batch axis; one all-reduce of the shared per-mode gradients.
* `rowpart`: one large mesh, contiguous node-row slabs of the block CSR per rank; the SpMM reads the
             halo rows of the dense block straight from the peers' memory over NVLink.
* `rowpart_lobpcg`: the eigen-solve itself on those slabs (peer-gather smoother, NCCL all-reduce of Gram strips / residual
             sums / partial coarse residuals, all-gather of the new search block, replicated small eigen-solve).
"""
from .sweep import (shard_indices, gather_ordered, sweep_modal_solves, WorkQueue, gather_indexed,  # noqa: F401
                    sweep_modal_solves_dynamic)
from .synth import batch_slice, sharded_modal_synth, allreduce_shared_grads  # noqa: F401
from .rowpart import slab_bounds, owner_of, RowPartition  # noqa: F401
from .rowpart_lobpcg import RowPartLOBPCG, eigen_decomposition_rowpart  # noqa: F401
