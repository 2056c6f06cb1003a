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
