"""Multi-GPU sharding of the transduction path.

The corpus is cut into byte-balanced, EOT-aligned shards, one per rank (one process
per GPU).  Every rank transduces its shard independently; the only exchange is an
all-gather of the per-shard counts {bytes, tokens, sentences, texts} and of the
carry (the walk state behind the shard's last EOT), from which every rank derives
the global index bases of its arrays.  Offsets themselves are text-relative, so
they need no fix-up.

A shard is walked from the guessed carry (root state, nothing pending).  The guess
is checked against the predecessor's real carry-out; on the rare mismatch (EOT
inside markup) the shard is transduced again from the true carry.
"""
import numpy as np

EOT = 4


def plan_shards(data: np.ndarray, n: int):
    """[(lo, hi)] * n: contiguous byte ranges, each ending right after an EOT (except the
    last, which ends at len(data)), as equal in size as the EOT positions allow."""
    N = int(data.size)
    if n <= 1 or N == 0:
        return [(0, N)] + [(N, N)] * (max(n, 1) - 1)
    cuts = [0]
    for r in range(1, n):
        target = max(cuts[-1], (N * r) // n)
        # first EOT at or after the balanced position
        window = 1 << 16
        cut = N
        p = target
        while p < N:
            idx = np.flatnonzero(data[p:min(N, p + window)] == EOT)
            if idx.size:
                cut = p + int(idx[0]) + 1
                break
            p += window
        cuts.append(cut)
    cuts.append(N)
    return [(cuts[i], cuts[i + 1]) for i in range(n)]


def exchange_counts(local_counts, group=None):
    """all-gather of this rank's [bytes, tokens, sentences, texts, sent_pos, carry_state] ->
    (per-rank matrix, exclusive bases of this rank).  Uses torch.distributed (NCCL over
    NVLink on GPUs, gloo on CPU); with no process group it is the identity."""
    import torch
    import torch.distributed as dist
    v = torch.as_tensor(local_counts, dtype=torch.int64)
    if not (dist.is_available() and dist.is_initialized()):
        return v.unsqueeze(0).numpy(), np.zeros_like(v.numpy())
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    mine = v.to(dev)
    parts = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    allc = torch.stack(parts).cpu().numpy()
    bases = allc[:rank].sum(axis=0) if rank else np.zeros(allc.shape[1], dtype=np.int64)
    return allc, bases


def carry_mismatch(all_counts, rank, root_state=1):
    """does the shard of `rank` have to be redone because the predecessor's walk did not end
    in the guessed state?  (column 5 = carry-out state of each shard)"""
    if rank == 0:
        return False
    prev = rank - 1
    while prev > 0 and all_counts[prev][0] == 0:  # empty shards pass the carry through
        prev -= 1
    if all_counts[prev][0] == 0:
        return False
    return int(all_counts[prev][5]) != root_state
