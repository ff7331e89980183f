"""Cross-page duplicate-figure removal: the path's single exchange step.

Pages shard across ranks with no data-path collective; each rank hashes its surviving regions
(64-bit integer-DCT pHash, csrc/phash.cu) and tags them with a global key (page_idx << 16 | region_idx).
ONE all-gather of fixed-capacity (hash, key) buffers (sentinel-padded, so no size exchange is needed)
gives every rank the whole corpus' hashes -- about 150 KB at 8,000 pages, latency-bound on NVLink --
and every rank then runs the same all-pairs Hamming kernel, so all ranks derive the identical survivor
set without a second collective.  The reference's nearest analogue is the id-dedup on md5 at
pdf_image_segmentation.py:3886-3887.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

SENTINEL_KEY = (1 << 63) - 1          # sorts after every real key (real keys are < 2^62)


def shard_pages(n_pages: int, rank: int, world: int) -> range:
    """Contiguous page block of `rank` (keeps page order trivial; sizes differ by at most one)."""
    base, rem = divmod(n_pages, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def region_key(page_idx: int, region_idx: int) -> int:
    if not (0 <= region_idx < 65536 and 0 <= page_idx < (1 << 45)):
        raise ValueError("region key out of range")
    return (page_idx << 16) | region_idx


def gather_hashes(hashes: torch.Tensor, keys: torch.Tensor, capacity: int, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-gather (hash, key) int64 pairs from every rank.  Works on CUDA tensors over NCCL and on CPU
    tensors over gloo (the host-logic tests).  Returns the concatenated valid (hashes, keys), sorted by key.

    Every intermediate has a FIXED shape (capacity x world): the sentinel-padded buffers are sorted by key, the
    sentinels sort last, and the result is a prefix view.  No data-dependent allocation means the caching allocator
    never has to cudaMalloc in the steady state -- with eight peer-mapped GPUs one cudaMalloc costs ~20 ms, which is
    what a boolean-mask compaction of the gathered buffer cost at N=8."""
    n = hashes.numel()
    if n > capacity:
        raise ValueError(f"rank holds {n} regions, more than the gather capacity {capacity}")
    buf = torch.full((capacity, 2), SENTINEL_KEY, dtype=torch.int64, device=hashes.device)
    buf[:n, 0] = hashes
    buf[:n, 1] = keys
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        world = dist.get_world_size(group)
        out = torch.empty((world * capacity, 2), dtype=torch.int64, device=hashes.device)
        dist.all_gather_into_tensor(out, buf, group=group)
    else:
        out = buf
    k_sorted, order = torch.sort(out[:, 1].contiguous())
    h_sorted = out[:, 0].contiguous()[order]
    n_valid = int((k_sorted != SENTINEL_KEY).sum().item())
    return h_sorted[:n_valid], k_sorted[:n_valid]


def init_comm(ctx, group=None) -> int:
    """Creates the library's own NCCL communicator over the ranks of the (already initialised) torch.distributed group:
    rank 0 draws the unique id (synseg_comm_unique_id), torch.distributed's store carries its 128 bytes to the other ranks,
    every rank calls synseg_comm_init.  Returns the world size (1: nothing to do)."""
    import ctypes as C
    from ._lib import check
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 1
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    buf = (C.c_uint8 * 128)()
    if rank == 0:
        check(ctx.lib.synseg_comm_unique_id(buf), "synseg_comm_unique_id")
    box = [bytes(buf)]
    dist.broadcast_object_list(box, src=0, group=group)
    ident = (C.c_uint8 * 128).from_buffer_copy(box[0])
    check(ctx.lib.synseg_comm_init(ctx._h, ident, rank, world), "synseg_comm_init")
    return world


def comm_info(ctx):
    import ctypes as C
    r, w, v = C.c_int32(), C.c_int32(), C.c_int32()
    ctx.lib.synseg_comm_info(ctx._h, C.byref(r), C.byref(w), C.byref(v))
    return dict(rank=r.value, world=w.value, nccl_version=v.value)


class DedupExchange:
    """Device buffers + the one call of the exchange step (synseg_dedup_exchange: pack, ONE ncclAllGather, replicated
    rank-and-dedup kernels).  `run` queues everything on the current stream and returns device tensors; nothing in it
    synchronises the host or launches a torch kernel."""

    def __init__(self, ctx, capacity: int, world: int = 1, max_hamming: int = 4):
        dev = ctx.device
        self.ctx, self.capacity, self.world, self.max_hamming = ctx, capacity, world, max_hamming
        total = capacity * world
        self.all_keys = torch.empty(total, dtype=torch.int64, device=dev)
        self.all_hashes = torch.empty(total, dtype=torch.int64, device=dev)
        self.keep = torch.empty(total, dtype=torch.uint8, device=dev)
        self.n_total = torch.zeros(1, dtype=torch.int32, device=dev)

    def run(self, hashes: torch.Tensor, keys: torch.Tensor, count: torch.Tensor):
        from ._lib import check
        c = self.ctx
        check(c.lib.synseg_dedup_exchange(c._h, hashes.data_ptr(), keys.data_ptr(), count.data_ptr(), self.capacity, self.max_hamming,
                                          self.all_keys.data_ptr(), self.all_hashes.data_ptr(), self.keep.data_ptr(), self.n_total.data_ptr(), c._s()),
              "synseg_dedup_exchange")
        return self.all_keys, self.keep, self.n_total

    def result(self):
        """(keys int64 [n] ascending, keep uint8 [n]) on the host -- the one synchronisation of the step."""
        n = int(self.n_total.item())
        return self.all_keys[:n].cpu(), self.keep[:n].cpu()


def survivors_digest(keys: torch.Tensor, keep: torch.Tensor) -> str:
    """sha1 over the surviving keys in ascending order: equal on every rank and for every world size over the same corpus."""
    import hashlib
    k = keys[keep.bool()].cpu().numpy().astype("<i8")
    return hashlib.sha1(k.tobytes()).hexdigest()[:16]


def cross_page_dedup(ctx, hashes: torch.Tensor, keys: torch.Tensor, capacity: int, max_hamming: int = 4, group=None, phases=None):
    """Returns (all_keys int64 [N] sorted, keep uint8 [N]) -- identical on every rank.  CUDA tensors go through the
    library's exchange (the communicator of `init_comm` when the job has several ranks); CPU tensors (the gloo host-logic
    tests) through `gather_hashes`."""
    if hashes.is_cuda:
        world = comm_info(ctx)["world"]
        ex = DedupExchange(ctx, capacity, world, max_hamming)
        count = torch.tensor([hashes.numel()], dtype=torch.int32, device=hashes.device)
        ex.run(hashes.contiguous(), keys.contiguous(), count)
        k, keep = ex.result()
        return k.to(hashes.device), keep.to(hashes.device)
    h, k = gather_hashes(hashes, keys, capacity, group)
    keep = ctx.phash_dedup(h, k, max_hamming)
    return k, keep
