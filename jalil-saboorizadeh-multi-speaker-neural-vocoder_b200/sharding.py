"""Utterance sharding for multi-GPU generation: one process per GPU, no data-path collective.

Utterances are independent (each owns its sample sequence, hidden states, conditioner and speaker), so rank r simply
generates the contiguous block `shard_range(n, r, world)`; results are gathered on the host by the caller."""


def shard_range(n_items, rank, world):
    """Contiguous, balanced partition of range(n_items): the first n_items % world ranks get one extra item."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_batch(cond, spk, uniforms, rank, world):
    """Slice per-utterance generation inputs: cond (B, n_cond, D), spk (B,), uniforms (T, B) -> this rank's block."""
    lo, hi = shard_range(cond.shape[0], rank, world)
    return cond[lo:hi], spk[lo:hi], uniforms[:, lo:hi], (lo, hi)
