"""Entity-partitioned TransE training across the GPUs of one NVLink box: host-side plumbing.

One process per GPU (torchrun).  The data path has no collective: the persistent kernels exchange rows
and updates with posted peer stores / REDs over NVLink (kb2e_b200/csrc/train_dist.cu).  torch.distributed is
used only to exchange the 64-byte CUDA IPC handles once and to add the per-rank losses."""
import numpy as np

from .api import Context, TABLE_ENTITY, TABLE_RELATION


def owned_ids(num_entities, rank, world):
    """Global ids of the entity rows rank owns, in local order (row e lives on rank e % world)."""
    return np.arange(rank, num_entities, world)


class PartitionedTrainer:
    def __init__(self, dim, num_entities, num_relations, rank, world, device, **cfg):
        self.rank, self.world = rank, world
        self.ctx = Context("transe", dim, num_entities, num_relations, device=device, **cfg)
        self._connected = False

    def set_training_set(self, train, head_mean, tail_mean):
        """Collective on first use: the batch size (len(train) // batches) sizes the exchange buffers of the
        arena, so the arenas are allocated and mapped into every peer once the training set is known."""
        self.ctx.set_train_triples(train)
        self.ctx.set_bern(head_mean, tail_mean)
        if self._connected:
            return
        import torch.distributed as dist
        handle = self.ctx.dist_setup(self.rank, self.world)
        if self.world > 1:
            handles = [None] * self.world
            dist.all_gather_object(handles, handle)
        else:
            handles = [handle]
        self.ctx.dist_connect(handles)
        self._connected = True
        if self.world > 1:
            dist.barrier()  # every arena is mapped everywhere before anyone launches

    def init_embeddings(self):
        self.ctx.dist_init_embeddings()
        self._sync()

    def upload_global(self, ent, rel):
        """Every rank passes the same global tables; each keeps its own entity rows."""
        self.ctx.dist_upload(TABLE_ENTITY, np.ascontiguousarray(ent[self.rank::self.world]))
        self.ctx.dist_upload(TABLE_RELATION, rel)
        self._sync()

    def train_epochs(self, first_epoch, n_epochs):
        """Collective.  Returns the global per-epoch loss."""
        err = None
        try:
            loss = self.ctx.dist_train_epochs(first_epoch, n_epochs)
        except Exception as e:   # raised on every rank below, after the ranks have agreed that the call failed
            err, loss = e, np.zeros(n_epochs)
        if self.world > 1:
            import torch
            import torch.distributed as dist
            t = torch.from_numpy(np.concatenate([loss, [1.0 if err is not None else 0.0]])).cuda()
            dist.all_reduce(t)   # one collective for the loss and the status: no rank is left waiting in it
            t = t.cpu().numpy()
            loss, failed = t[:-1], t[-1] > 0
            if failed and err is None:
                err = RuntimeError("partitioned training failed on another rank")
        if err is not None:
            raise err
        return loss

    def gather_global(self):
        """Global (entity, relation) tables on every rank (for evaluation / writing files)."""
        local = self.ctx.dist_download(TABLE_ENTITY)
        rel = self.ctx.dist_download(TABLE_RELATION)
        if self.world == 1:
            return local, rel
        import torch.distributed as dist
        parts = [None] * self.world
        dist.all_gather_object(parts, local)
        ent = np.empty((self.ctx.nE, self.ctx.dim), dtype=np.float64)
        for g, p in enumerate(parts):
            ent[g::self.world] = p
        return ent, rel

    def _sync(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()

    def close(self):
        self._sync()
        self.ctx.close()
