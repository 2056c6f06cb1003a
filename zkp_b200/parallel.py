"""Multi-GPU plumbing for the one exchange the path has: independent proofs shard across
ranks with NO data-path collective; each rank batch-verifies its own contiguous proof range (with its own random
weights, so every shard is a complete BatchVerifier::verify_batchable check,
/root/reference/src/toolbox/batch_verifier.rs:137-235) and the ranks exchange one accept bit each in a single
all-gather (NCCL over NVLink on GPUs; gloo in the CPU tests).  The batch verdict is the AND of the bits."""
import torch
import torch.distributed as dist


def shard_range(n_proofs, rank, world):
    """Contiguous proof range [lo, hi) of `rank`; ranges differ by at most one proof."""
    per, rem = divmod(n_proofs, world)
    lo = rank * per + min(rank, rem)
    return lo, lo + per + (1 if rank < rem else 0)


def shard_columns(rows_by_proof, rank, world):
    """Slice instance rows shaped (rows, N, 32) to this rank's proofs: the reference's point/coefficient order is
    row-major [row][proof] (batch_verifier.rs:208-222), so a proof shard is a column slab of every row."""
    lo, hi = shard_range(rows_by_proof.shape[1], rank, world)
    return rows_by_proof[:, lo:hi]


def gather_accept_bits(accept, device=None, group=None):
    """All-gather one accept bit per rank; returns (bits list in rank order, overall verdict)."""
    if not (dist.is_available() and dist.is_initialized()):
        return [int(bool(accept))], bool(accept)
    world = dist.get_world_size(group)
    mine = torch.tensor([1 if accept else 0], dtype=torch.int32, device=device)
    out = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    bits = [int(t.item()) for t in out]
    return bits, all(bits)


def gather_partial_sums(partial_limbs, device=None, group=None):
    """Single-verdict mode: all-gather every rank's partial MSM sum -- one extended point in limb form,
    20 x uint64 = 160 bytes per rank -- and return the (world, 20) array every rank feeds to Engine.partials_verdict.
    A rank whose shard held an invalid encoding passes None: the batch is rejected (VerificationFailure)."""
    import numpy as np
    ok = partial_limbs is not None
    mine = np.zeros(21, dtype=np.int64)
    if ok:
        mine[:20] = np.ascontiguousarray(partial_limbs, dtype=np.uint64).reshape(20).view(np.int64)
        mine[20] = 1
    if not (dist.is_available() and dist.is_initialized()):
        return (mine[:20].view(np.uint64).reshape(1, 20) if ok else None)
    world = dist.get_world_size(group)
    t = torch.from_numpy(mine).to(device) if device is not None else torch.from_numpy(mine)
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    arr = np.stack([o.cpu().numpy() for o in out])
    if not (arr[:, 20] == 1).all():
        return None
    return np.ascontiguousarray(arr[:, :20]).view(np.uint64)
