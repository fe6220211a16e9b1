"""Placement of Paxos groups over the GPUs of one box and the peer-handle exchange
between the per-GPU processes (the stand-in for APUS's RC_SYN handshake,
dare_ibv_ud.c:1168-1416).  Pure host logic, no CUDA: covered on CPU with gloo.

A group does not shard; scale-out = independent groups.  With one
process per GPU, group g is led by rank g and its replica r lives on rank (g + r) % N,
so every GPU leads one group and follows in `replicas - 1` others (weak scaling).
"""


def hosted(rank: int, world: int, replicas: int):
    """[(group, replica_idx)] hosted by `rank`."""
    return [(g, r) for g in range(world) for r in range(replicas) if (g + r) % world == rank]


def host_of(group: int, replica: int, world: int) -> int:
    return (group + replica) % world


def exchange(dist, local_blobs: dict, world: int, group=None) -> dict:
    """all-gather {(group, replica): 128-byte handle} over torch.distributed."""
    gathered = [None] * world
    dist.all_gather_object(gathered, local_blobs, group=group)
    merged = {}
    for d in gathered:
        for k, v in d.items():
            assert k not in merged, f"replica {k} exported twice"
            merged[k] = v
    return merged


def connections(rank: int, world: int, replicas: int):
    """[(group, replica_idx, peer_idx)]: which peer handles each local replica must map."""
    return [(g, r, p) for (g, r) in hosted(rank, world, replicas) for p in range(replicas) if p != r]
