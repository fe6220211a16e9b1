"""A model of the leader's space / pruning arithmetic (apus_kernels.cu: leader_place -- ring offsets, head := the
furthest-behind apply offset, free-space rule E2) checked against ground truth kept in ABSOLUTE byte positions.

It documents the round-2 finding behind DESIGN.md section 3a: apply offsets are ring offsets, so a SNAPSHOT of them that is
used after other workers have pruned can alias into the used region of the next lap ("almost caught up") and let the head
overtake an application that is a lap behind -- the model reproduces that within a few dozen tiles -- while offsets read
at the time of use (what the kernel does now: under the place turn) never do.  Pure Python, no GPU."""
import random

import pytest

L = 1 << 20
ES, TILE = 264, 256


def rd(a, b):
    return b - a if b >= a else L - (a - b)


def run(seed, stale_snapshots, n_tiles=300):
    rnd = random.Random(seed)
    pos_of = {0: 0}                                   # absolute entry boundary -> ring offset
    end_abs = head_abs = 0
    end, head, tail, prev_head, was_blocked = L, 0, 0, False, False
    true_ap = [0, 0]                                  # the followers' applications, absolute (ground truth)
    hist = [[(0, 0)], [(0, 0)]]                       # what they forwarded to the leader: (absolute, ring offset)
    tiles = guard = 0
    while tiles < n_tiles:
        guard += 1
        assert guard < 200000
        for f in range(2):
            if rnd.random() < 0.7:
                true_ap[f] = min(end_abs, true_ap[f] + rnd.randrange(0, 200000))
            if rnd.random() < 0.5 and true_ap[f] in pos_of:
                hist[f].append((true_ap[f], pos_of[true_ap[f]]))
        aps = [(end_abs, 0 if end == L else end)]     # the leader's own apply == its commit
        for f in range(2):
            k = rnd.randrange(0, 3) if stale_snapshots else 0
            aps.append(hist[f][max(0, len(hist[f]) - 1 - k)])
        pos0 = 0 if end == L else end
        used = 0 if end == L else rd(head, end)
        autoh = False
        if end != L and used >= (L >> 2) and (not prev_head or was_blocked) and L - pos0 >= 64:
            d = 0
            for _, ring in aps:
                d = max(d, min(rd(ring, end), used))
            if d == 0:
                d = rd(tail, end)
            if d <= used and used - d >= (L >> 3):
                autoh, used = True, d
                head = end - d if end >= d else L - (d - end)
                head_abs = end_abs - d
                for f in range(2):
                    if head_abs > true_ap[f]:
                        return f"head passed follower {f}'s application by {head_abs - true_ap[f]} bytes after {tiles} tiles"
        hbytes = 64 if autoh else 0
        lim_space = L - used - 1 - 64 if L - used > 65 else 0
        limit = min(L - pos0, lim_space)
        m = 0
        while m < TILE and hbytes + (m + 1) * ES <= limit:
            m += 1
        if m == 0 and not autoh:
            left = L - pos0
            if ES > left and used + left + ES + 64 < L:
                end_abs += left
                end = 0
                continue
            was_blocked = True
            continue
        was_blocked = False
        p_abs, p = end_abs, pos0
        if autoh:
            pos_of[p_abs] = p; p_abs += 64; p += 64
        for _ in range(m):
            pos_of[p_abs] = p; p_abs += ES; p += ES
        pos_of[p_abs] = 0 if p == L else p
        tail = p - (ES if m else 64)
        end_abs, end = p_abs, (0 if p == L else p)
        prev_head = autoh and m == 0
        assert end_abs - head_abs < L, "ring overfull"
        tiles += 1
    return None


def test_stale_apply_snapshots_alias_and_overtake_an_application():
    hits = [run(seed, stale_snapshots=True) for seed in range(40)]
    assert any(hits), "the model no longer reproduces the aliasing it documents"


@pytest.mark.parametrize("block", range(4))
def test_apply_offsets_read_at_time_of_use_never_overtake(block):
    for seed in range(block * 150, (block + 1) * 150):
        assert run(seed, stale_snapshots=False) is None
