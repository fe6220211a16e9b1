/*
 * apus_kernels.cu -- the persistent sm_100a kernels of the replication engine.
 *
 * One kernel, `apus_replica_kernel`, launched with one CTA per replica ROLE that
 * lives on the launching GPU (apus_role_t table).  Two roles:
 *
 *   LEADER   the hot loop of the reference's leader, fused:
 *            get_tailq_message (dare_ibv_ud.c:780-790) + log_append_entry
 *            (dare_log.h:466-558) + persist_new_entries' sender stamp
 *            (dare_server.c:1803-1804) + update_remote_logs step I/II
 *            (dare_ibv_rc.c:1526-1573: byte range then tail) + the commit rule
 *            (dare_ibv_rc.c:1725-1758) + the commit publish (:1760-1822).
 *            15 producer warps build a TILE of entries in shared memory and push
 *            it with 16 B vector stores into the local log and into every
 *            follower's log over NVLink; warp 15 is the commit warp: lane i
 *            polls follower i's ack word and a shuffle/ballot ranks the acks to
 *            find the offset a majority holds.
 *   FOLLOWER persist_new_entries' follower branch (dare_server.c:1792-1810) +
 *            rc_send_entries_reply (dare_ibv_rc.c:1828-1863): poll `end`, walk
 *            the new entries, set reply[me] locally and in the leader's copy,
 *            publish the ack word, follow `commit`.
 *
 * Ordering (invariant I1, "data before tail"): all data stores of a tile ->
 * bar.sync -> fence.acq_rel.sys -> st.relaxed.sys of `end`.  The follower reads
 * `end` with ld.acquire.sys and the entry bytes with ld.relaxed.sys (never
 * through a stale L1 line).  Acks mirror this in the other direction.
 *
 * Pure integer / byte work: no tensor cores, bound by NVLink store bandwidth and
 * by launch-free round-trip latency.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "apus_layout.h"

// ---------------------------------------------------------------------------------
// memory-model helpers (system scope: peers and the host observe these)
// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t ld_relaxed_sys(const volatile void *p)
{
    uint64_t v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint64_t ld_acquire_sys(const volatile void *p)
{
    uint64_t v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_relaxed_sys_u32(const volatile void *p)
{
    uint32_t v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys(volatile void *p, uint64_t v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys_u8(volatile void *p, uint8_t v)
{
    asm volatile("st.relaxed.sys.global.u8 [%0], %1;" ::"l"(p), "r"((uint32_t)v) : "memory");
}
__device__ __forceinline__ uint4 ld_relaxed_sys_v4(const void *p)
{
    uint4 v;
    asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p)
                 : "memory");
    return v;
}
__device__ __forceinline__ void st_v4(void *p, uint4 v)
{
    asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ void st_u8(void *p, uint32_t v)
{
    asm volatile("st.global.u8 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint64_t globaltimer_ns()
{
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// named barrier for a subset of the CTA's warps
__device__ __forceinline__ void bar_sync(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ bool has_cmd(uint32_t type)
{
    return !(type == T_NOOP || type == T_CONFIG || type == T_HEAD);
}
// bytes placed at entry+48 (sm_cmd_t image {u16 len; cmd}, dare_cid_t, or head offset)
__device__ __forceinline__ uint32_t data_bytes(uint32_t type, uint32_t len)
{
    if (type == T_NOOP) return 0;
    if (type == T_CONFIG) return 16;
    if (type == T_HEAD) return 8;
    return 2u + len;
}
__device__ __forceinline__ uint32_t entry_stride(uint32_t type, uint32_t len)
{
    return has_cmd(type) ? APUS_HDR_BYTES + len : APUS_HDR_BYTES;   // dare_log.h:228-234
}

#define WATCHDOG_NS (20ull * 1000ull * 1000ull * 1000ull)

// error codes reported through hostwords.error
#define APUS_KERR_WATCHDOG_LEADER   1
#define APUS_KERR_WATCHDOG_FOLLOWER 2
#define APUS_KERR_WATCHDOG_COMMIT   3
#define APUS_KERR_BAD_ENTRY         4

// ---------------------------------------------------------------------------------
// shared memory
// ---------------------------------------------------------------------------------
#define PUB_RING 256u
#define N_PRODUCER_WARPS 15
#define N_PRODUCER_THREADS (N_PRODUCER_WARPS * 32)

struct LeaderShared {
    // tile table (one row per entry of the tile)
    uint64_t req_id[APUS_MAX_TILE_ENTRIES];
    uint32_t pay_off[APUS_MAX_TILE_ENTRIES];   // 16 B units into the payload ring
    uint32_t rel[APUS_MAX_TILE_ENTRIES];       // entry start - tile start (bytes)
    uint16_t len[APUS_MAX_TILE_ENTRIES];
    uint16_t clt[APUS_MAX_TILE_ENTRIES];
    uint8_t  type[APUS_MAX_TILE_ENTRIES];
    // tile control (written by thread 0 / warp 0, read by all producers)
    uint32_t n_fetch;        // descriptors fetched this round
    uint32_t m;              // entries in the tile
    uint32_t gap;            // 1: wrap-gap tile (range [a, len), optional ghost of entry 0)
    uint32_t ghost;          // 1: ghost header is composed at a
    uint32_t fresh;          // 1: the range was never written (holes are zero)
    uint32_t finish;         // producers are done
    uint32_t auto_head;      // 1: the tile starts with a HEAD entry appended by the pruning rule
    uint64_t auto_head_val;  // ... carrying this head offset
    uint64_t a, b;           // byte range of the tile in the log
    uint64_t idx0;           // idx of the tile's first entry
    uint64_t t_dequeue;
    uint8_t *peer_entries[APUS_MAX_SERVERS];
    // publishes in flight: producer -> commit warp
    uint64_t pub_cum[PUB_RING];      // entries published up to and including this tile
    uint64_t pub_end[PUB_RING];      // `end` after this tile
    uint64_t pub_tickets[PUB_RING];  // tickets consumed up to and including this tile
    uint64_t pub_t0[PUB_RING];
    volatile uint64_t pub_head;      // next slot the producer writes
    volatile uint64_t pub_tail;      // next slot the commit warp reads
    volatile uint64_t published;     // entries published (the leader's own "ack")
    volatile uint32_t producers_done;
    volatile uint32_t abort_flag;
};

struct FollowerShared {
    uint32_t off[APUS_IMG_BYTES / 64 + 8];   // entry offsets found in the window (log offsets, low 32 bits)
    uint32_t n;
    uint32_t done;
    uint64_t win_lo, win_hi, next;           // window bounds in the log, next walk offset
    uint64_t head_val, head_end;             // last HEAD entry of the window (head_end == len: none)
    uint64_t end_seen, commit_seen;
};

extern __shared__ __align__(16) uint8_t smem_raw[];

// ---------------------------------------------------------------------------------
// LEADER
// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t hdr_byte(uint32_t j, uint64_t idx, uint64_t term, uint64_t req_id,
                                             uint32_t clt, uint32_t type, uint32_t sender)
{
    // dare_log_entry_t bytes 0..40 (dare_log.h:33-48)
    if (j < 8) return (uint32_t)(idx >> (8 * j)) & 0xFF;
    if (j < 16) return (uint32_t)(term >> (8 * (j - 8))) & 0xFF;
    if (j < 24) return (uint32_t)(req_id >> (8 * (j - 16))) & 0xFF;
    if (j == 24) return clt & 0xFF;
    if (j == 25) return (clt >> 8) & 0xFF;
    if (j == 26) return type;
    if (j == 27) return sender;
    return 0;   // reply[13]
}

// copy nbytes from a 16 B-aligned global source to an arbitrarily aligned smem destination
__device__ __forceinline__ void warp_copy_to_smem(uint8_t *dst, const uint8_t *src, uint32_t nbytes, int lane)
{
    uint32_t nchunks = (nbytes + 15u) >> 4;
    uint32_t dalign = (uint32_t)(uintptr_t)dst & 15u;
    for (uint32_t c = lane; c < nchunks; c += 32) {
        uint4 v = ld_relaxed_sys_v4(src + 16u * c);
        uint8_t *d = dst + 16u * c;
        uint32_t left = nbytes - 16u * c;
        if (dalign == 0 && left >= 16) {
            *reinterpret_cast<uint4 *>(d) = v;
        } else if ((dalign & 3u) == 0 && left >= 16) {
            uint32_t *d4 = reinterpret_cast<uint32_t *>(d);
            d4[0] = v.x; d4[1] = v.y; d4[2] = v.z; d4[3] = v.w;
        } else {
            uint32_t w[4] = {v.x, v.y, v.z, v.w};
            uint32_t nb = left < 16 ? left : 16;
#pragma unroll
            for (uint32_t k = 0; k < 16; k++)
                if (k < nb) d[k] = (uint8_t)(w[k >> 2] >> (8 * (k & 3)));
        }
    }
}

__device__ void leader_commit_warp(const apus_devctx_t *__restrict__ cx, LeaderShared *S)
{
    const int lane = threadIdx.x & 31;
    const int N = cx->group_size, me = cx->idx, quorum = cx->quorum;
    apus_ctrl_t *ctrl = reinterpret_cast<apus_ctrl_t *>(cx->region);
    apus_loghdr_t *hdr = reinterpret_cast<apus_loghdr_t *>(cx->region + APUS_CTRL_BYTES);
    apus_hostwords_t *hw = cx->hw;
    uint64_t committed = ctrl->committed;
    uint64_t committed_tickets = ctrl->committed_tickets;
    uint64_t lat_count = ctrl->lat_count;
    uint64_t last_progress = globaltimer_ns();
    uint32_t spins = 0;
    volatile uint64_t *peer_commit = nullptr;
    if (lane < N && lane != me && cx->peer[lane])
        peer_commit = &reinterpret_cast<apus_loghdr_t *>(cx->peer[lane] + APUS_CTRL_BYTES)->commit;

    for (;;) {
        // lane i holds what replica i has acked (entries, monotone); the leader's own
        // vote is everything it has published (dare_ibv_rc.c:1736 "i == idx")
        uint64_t v = 0;
        if (lane < N) v = (lane == me) ? S->published : ld_relaxed_sys(&ctrl->ack[lane]);
        // rank: how many replicas hold at least what I hold
        int cnt = 0;
        for (int j = 0; j < N; j++) {
            uint64_t vj = __shfl_sync(0xffffffffu, v, j);
            cnt += (vj >= v) ? 1 : 0;
        }
        uint64_t cand = (lane < N && cnt >= quorum) ? v : 0;
        // the largest count a majority holds (size/2+1, dare_ibv_rc.c:1741)
        for (int s = 16; s > 0; s >>= 1) {
            uint64_t o = __shfl_xor_sync(0xffffffffu, cand, s);
            cand = o > cand ? o : cand;
        }
        const uint64_t Q = cand;
        if (Q > committed) {
            __threadfence_system();   // acquire side of the followers' ack publication
            // map the entry count to the log offset recorded at publish time
            uint64_t tail = S->pub_tail, head = S->pub_head;
            uint64_t off = 0, tickets = committed_tickets, t0 = 0;
            bool any = false;
            while (tail != head && S->pub_cum[tail & (PUB_RING - 1)] <= Q) {
                off = S->pub_end[tail & (PUB_RING - 1)];
                tickets = S->pub_tickets[tail & (PUB_RING - 1)];
                t0 = S->pub_t0[tail & (PUB_RING - 1)];
                if ((cx->flags & 0x2u) && cx->lat_ns && lane == 0) {
                    uint64_t d = globaltimer_ns() - t0;
                    cx->lat_ns[lat_count & (APUS_LAT_RING - 1)] = d > 0xffffffffull ? 0xffffffffu : (uint32_t)d;
                }
                lat_count++;
                committed = S->pub_cum[tail & (PUB_RING - 1)];
                tail++;
                any = true;
            }
            if (any) {
                // commit is a prefix and an entry boundary (invariant I3)
                if (peer_commit) st_relaxed_sys(peer_commit, off);          // dare_ibv_rc.c:1810
                if (lane == 0) {
                    hdr->commit = off;
                    hdr->apply = off;                                       // leader applies = update_state
                    ctrl->committed = committed;
                    ctrl->committed_tickets = tickets;
                    ctrl->lat_count = lat_count;
                    __threadfence_system();
                    st_relaxed_sys(&hw->commit_off, off);
                    st_relaxed_sys(&hw->committed_tickets, tickets);        // releases proxy.c:160 spinners
                    S->pub_tail = tail;
                }
                committed_tickets = tickets;
                last_progress = globaltimer_ns();
                __syncwarp();
            }
        }
        // exit: producers finished and nothing is in flight
        int ex = 0;
        if (lane == 0) {
            if (S->producers_done && committed == S->published) ex = 1;
            else if ((++spins & 0x3ffu) == 0) {
                if (ld_relaxed_sys_u32(&hw->stop) || S->abort_flag) ex = 2;
                else if (cx->target != ~0ull && globaltimer_ns() - last_progress > WATCHDOG_NS &&
                         committed != S->published) {
                    st_relaxed_sys(&hw->error, APUS_KERR_WATCHDOG_COMMIT);
                    S->abort_flag = 1;
                    ex = 2;
                }
            }
        }
        ex = __shfl_sync(0xffffffffu, ex, 0);
        if (ex) {
            // clean end of a bounded launch: tell every follower how many entries exist, so
            // that it can leave once it has acked and applied all of them
            if (ex == 1 && cx->target != ~0ull && lane < N && lane != me && cx->peer[lane]) {
                apus_ctrl_t *pc = reinterpret_cast<apus_ctrl_t *>(cx->peer[lane]);
                st_relaxed_sys(&pc->fin_entries, committed);
                __threadfence_system();
                st_relaxed_sys(&pc->fin_target, cx->target);
            }
            break;
        }
    }
}

__device__ void leader_main(const apus_devctx_t *__restrict__ cx)
{
    LeaderShared *S = reinterpret_cast<LeaderShared *>(smem_raw);
    uint8_t *img = smem_raw + ((sizeof(LeaderShared) + 127u) & ~127u);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = cx->group_size, me = cx->idx;
    apus_ctrl_t *ctrl = reinterpret_cast<apus_ctrl_t *>(cx->region);
    apus_loghdr_t *hdr = reinterpret_cast<apus_loghdr_t *>(cx->region + APUS_CTRL_BYTES);
    uint8_t *entries = cx->region + APUS_ENTRIES_OFF;
    apus_hostwords_t *hw = cx->hw;
    const uint64_t L = cx->log_len;

    if (tid == 0) {
        S->pub_head = 0; S->pub_tail = 0;
        S->published = ctrl->published;
        S->producers_done = 0; S->abort_flag = 0; S->finish = 0;
        for (int i = 0; i < APUS_MAX_SERVERS; i++)
            S->peer_entries[i] = (i < N && i != me && cx->peer[i]) ? cx->peer[i] + APUS_ENTRIES_OFF : nullptr;
        // entries published by an earlier launch but not yet committed come back as one record
        if (ctrl->published != ctrl->committed) {
            S->pub_cum[0] = ctrl->published; S->pub_end[0] = hdr->end;
            S->pub_tickets[0] = ctrl->consumed; S->pub_t0[0] = globaltimer_ns();
            S->pub_head = 1;
        }
    }
    __syncthreads();

    if (warp == N_PRODUCER_WARPS) {   // warp 15
        leader_commit_warp(cx, S);
        return;
    }

    // ---- producer warps 0..14 ------------------------------------------------------
    // thread-0 state mirrored in registers of every producer thread after each barrier
    uint64_t end = hdr->end, tailpos = hdr->tail, next_idx = ctrl->next_idx;
    uint64_t consumed = ctrl->consumed, published = ctrl->published, hwm = ctrl->hwm;
    uint64_t bytes_rep = ctrl->bytes_replicated, batches = ctrl->batches;
    uint64_t auto_heads = ctrl->auto_heads;
    uint64_t last_progress = globaltimer_ns();
    bool pending_gap = false;   // a gap tile was stored and awaits the next publish
    bool prev_head = false;     // the last entry appended is a HEAD entry of the pruning rule

    for (;;) {
        // ---- T0: wait for requests (thread 0) ---------------------------------------
        if (tid == 0) {
            uint32_t n = 0, fin = 0, spins = 0;
            for (;;) {
                if (consumed >= cx->target) { fin = 1; break; }
                uint64_t t = ld_relaxed_sys(cx->sub_tail);
                uint64_t avail = t - consumed;
                if (avail) {
                    uint64_t room = cx->target - consumed;
                    if (avail > room) avail = room;
                    n = avail > APUS_MAX_TILE_ENTRIES ? APUS_MAX_TILE_ENTRIES : (uint32_t)avail;
                    // do not overrun the in-flight publish ring
                    uint32_t w = 0;
                    while (S->pub_head - S->pub_tail >= PUB_RING - 2) {
                        if (S->abort_flag || ((++w & 0xfffu) == 0 && ld_relaxed_sys_u32(&hw->stop))) {
                            n = 0; fin = 1; break;
                        }
                    }
                    break;
                }
                if ((++spins & 0xffu) == 0) {
                    if (ld_relaxed_sys_u32(&hw->stop) || S->abort_flag) { fin = 1; break; }
                    if (cx->target != ~0ull && globaltimer_ns() - last_progress > WATCHDOG_NS) {
                        st_relaxed_sys(&hw->error, APUS_KERR_WATCHDOG_LEADER);
                        S->abort_flag = 1; fin = 1; break;
                    }
                }
            }
            if (n) __threadfence_system();   // acquire: descriptors + payload behind the doorbell
            S->n_fetch = n; S->finish = fin;
            S->t_dequeue = globaltimer_ns();
        }
        bar_sync(1, N_PRODUCER_THREADS);
        if (S->finish) break;
        const uint32_t nf = S->n_fetch;

        // ---- T1: fetch descriptors (coalesced 16 B loads) ----------------------------
        for (uint32_t k = tid; k < nf; k += N_PRODUCER_THREADS) {
            uint4 d = ld_relaxed_sys_v4(&cx->sub_desc[(consumed + k) & cx->sub_mask]);
            S->req_id[k] = (uint64_t)d.x | ((uint64_t)d.y << 32);
            S->type[k] = (uint8_t)(d.z >> 24);
            S->pay_off[k] = d.z & 0x00ffffffu;
            S->len[k] = (uint16_t)(d.w & 0xffffu);
            S->clt[k] = (uint16_t)(d.w >> 16);
        }
        bar_sync(1, N_PRODUCER_THREADS);

        // ---- T2: placement (warp 0): log_append_entry's offset rules over the tile ----
        if (warp == 0) {
            uint64_t head = ld_relaxed_sys(&hdr->head);
            const uint64_t pos0 = (end == L) ? 0 : end;               // empty log starts at 0 (dare_log.h:216-219)
            uint64_t used = (end == L) ? 0 : (end >= head ? end - head : L - (head - end));
            // ---- device-side log pruning (log_pruning / force_log_pruning,
            //      dare_server.c:1996-2122): head := the smallest apply offset in the group,
            //      published through a HEAD entry placed in front of this tile
            uint32_t autoh = 0;
            uint64_t new_head = 0;
            if ((cx->flags & APUS_FLAG_AUTOPRUNE) && end != L && used >= (L >> 2) && !prev_head &&
                L - pos0 >= APUS_HDR_BYTES) {
                uint64_t d = 0;                                   // distance apply -> end, per replica
                if (lane < N) {
                    const uint64_t ap = (lane == me) ? ld_relaxed_sys(&hdr->apply) : ld_relaxed_sys(&ctrl->apply_off[lane]);
                    d = (end >= ap) ? end - ap : L - (ap - end);
                    if (d > used) d = used;                       // never behind the current head
                }
                for (int sft = 16; sft > 0; sft >>= 1) {
                    const uint64_t o = __shfl_xor_sync(0xffffffffu, d, sft);
                    d = o > d ? o : d;
                }
                if (d == 0) d = (end >= tailpos) ? end - tailpos : L - (tailpos - end);   // leave one entry (:2031-2034)
                if (d <= used && used - d >= (L >> 3)) {
                    autoh = 1;
                    new_head = (end >= d) ? end - d : L - (d - end);
                    used = d;                                     // the head moves before the append (:2041)
                    head = new_head;
                }
            }
            const uint64_t hbytes = autoh ? APUS_HDR_BYTES : 0;
            // limits for a contiguous tile starting at pos0
            uint64_t lim = L - pos0;                                   // no entry may cross len
            const uint64_t imgcap = APUS_IMG_BYTES - 16u - (pos0 & 15u);
            if (lim > imgcap) lim = imgcap;
            // rule E2: stay strictly before head (keep room for one HEAD entry when pruning on the device)
            const uint64_t reserve = (cx->flags & APUS_FLAG_AUTOPRUNE) ? APUS_HDR_BYTES : 0;
            const uint64_t lim_space = (L - used > 1 + reserve) ? (L - used - 1 - reserve) : 0;
            // per-lane strip of entries, two-level exclusive scan of strides
            const uint32_t per = (nf + 31u) / 32u;
            const uint32_t k0 = lane * per, k1 = (k0 + per < nf) ? k0 + per : nf;
            uint32_t sum = 0;
            for (uint32_t k = k0; k < k1; k++) sum += entry_stride(S->type[k], S->len[k]);
            uint32_t incl = sum;
            for (int sft = 1; sft < 32; sft <<= 1) {
                uint32_t o = __shfl_up_sync(0xffffffffu, incl, sft);
                if (lane >= sft) incl += o;
            }
            uint64_t run = hbytes + (incl - sum);   // bytes before my strip (behind the optional HEAD entry)
            uint32_t first_bad = nf;
            for (uint32_t k = k0; k < k1; k++) {
                uint32_t es = entry_stride(S->type[k], S->len[k]);
                S->rel[k] = (uint32_t)run;
                if (first_bad == nf && (run + es > lim || run + es > lim_space)) first_bad = k;
                run += es;
            }
            for (int sft = 16; sft > 0; sft >>= 1) {
                uint32_t o = __shfl_xor_sync(0xffffffffu, first_bad, sft);
                first_bad = o < first_bad ? o : first_bad;
            }
            if (lane == 0) {
                uint32_t m = first_bad;
                S->gap = 0; S->ghost = 0;
                if (m == 0 && !autoh) {
                    // entry 0 does not fit at pos0: wrap (dare_log.h:502-504, 526-538) or no space
                    const uint32_t es0 = entry_stride(S->type[0], S->len[0]);
                    const uint64_t left = L - pos0;
                    if (es0 > left && used + left + es0 + reserve < L) {
                        S->gap = 1;
                        S->ghost = (left >= APUS_HDR_BYTES && has_cmd(S->type[0])) ? 1u : 0u;   // header fits: ghost stays behind
                        S->a = pos0; S->b = L;
                    } else {
                        S->a = S->b = pos0;   // back-pressure: wait for head to advance
                    }
                } else {
                    S->a = pos0;
                    S->b = pos0 + (m ? S->rel[m - 1] + entry_stride(S->type[m - 1], S->len[m - 1]) : hbytes);
                }
                S->m = m;
                S->auto_head = autoh; S->auto_head_val = new_head;
                if (autoh) st_relaxed_sys(&hdr->head, new_head);
                S->idx0 = next_idx;
                S->fresh = (S->a >= hwm) ? 1u : 0u;
            }
        }
        bar_sync(1, N_PRODUCER_THREADS);
        const uint32_t m = S->m, gap = S->gap;
        const uint64_t a = S->a, b = S->b;
        if (a == b) {   // no space before head: poll again (stop / watchdog handled in T0)
            if (tid == 0) {
                if (ld_relaxed_sys_u32(&hw->stop) || S->abort_flag) S->finish = 1;
                else if (cx->target != ~0ull && globaltimer_ns() - last_progress > WATCHDOG_NS) {
                    st_relaxed_sys(&hw->error, APUS_KERR_WATCHDOG_LEADER);
                    S->abort_flag = 1; S->finish = 1;
                }
            }
            bar_sync(1, N_PRODUCER_THREADS);
            if (S->finish) break;
            continue;
        }
        const uint64_t a16 = a & ~15ull;
        const uint32_t nchunks = (uint32_t)(((b + 15ull) & ~15ull) - a16) >> 4;

        // ---- T3: prefill the image: zeros when the range is fresh, else the bytes the
        //      local log holds (holes of an entry keep what was there, like the reference)
        if (S->fresh) {
            for (uint32_t c = tid; c < nchunks; c += N_PRODUCER_THREADS)
                reinterpret_cast<uint4 *>(img)[c] = make_uint4(0, 0, 0, 0);
        } else {
            for (uint32_t c = tid; c < nchunks; c += N_PRODUCER_THREADS)
                reinterpret_cast<uint4 *>(img)[c] = ld_relaxed_sys_v4(entries + a16 + 16ull * c);
        }
        bar_sync(1, N_PRODUCER_THREADS);

        // ---- T4: compose entries into the image -----------------------------------------
        if (gap) {
            if (S->ghost && warp == 0) {
                // header of entry 0 without payload, sender untouched (dare_log.h:496-503, 521)
                uint8_t *e = img + (a - a16);
                const uint32_t ty = S->type[0];
                for (uint32_t j = lane; j < 41; j += 32)
                    if (j != E_SENDER)
                        e[j] = (uint8_t)hdr_byte(j, next_idx, cx->term, S->req_id[0], S->clt[0], ty, 0);
                if (lane == 0) { e[E_DATA] = (uint8_t)(S->len[0] & 0xff); e[E_DATA + 1] = (uint8_t)(S->len[0] >> 8); }
            }
        } else {
            const uint32_t autoh = S->auto_head;
            if (autoh && warp == N_PRODUCER_WARPS - 1) {
                // <HEAD, head_offset> entry (dare_log.h:29-32, dare_server.c:2043-2046)
                uint8_t *e = img + (a - a16);
                for (uint32_t j = lane; j < 41; j += 32)
                    e[j] = (uint8_t)hdr_byte(j, S->idx0, cx->term, 0, 0, T_HEAD, me);
                if (lane < 8) e[E_DATA + lane] = (uint8_t)(S->auto_head_val >> (8 * lane));
            }
            for (uint32_t k = warp; k < m; k += N_PRODUCER_WARPS) {
                uint8_t *e = img + (a - a16) + S->rel[k];
                const uint32_t ty = S->type[k], ln = S->len[k];
                const uint64_t rq = S->req_id[k];
                const uint32_t cl = S->clt[k];
                for (uint32_t j = lane; j < 41; j += 32)
                    e[j] = (uint8_t)hdr_byte(j, S->idx0 + autoh + k, cx->term, rq, cl, ty, me);
                const uint32_t nb = data_bytes(ty, ln);
                if (nb) warp_copy_to_smem(e + E_DATA, cx->sub_pay + 16ull * S->pay_off[k], nb, lane);
            }
        }
        bar_sync(1, N_PRODUCER_THREADS);

        // ---- T5: push the byte range [a,b) to the local log and to every follower ------
        for (uint32_t c = tid; c < nchunks; c += N_PRODUCER_THREADS) {
            const uint64_t lo = a16 + 16ull * c;
            const uint4 v = reinterpret_cast<const uint4 *>(img)[c];
            if (lo >= a && lo + 16 <= b) {
                st_v4(entries + lo, v);
#pragma unroll 1
                for (int f = 0; f < N; f++)
                    if (S->peer_entries[f]) st_v4(S->peer_entries[f] + lo, v);
            } else {
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
                for (uint32_t j = 0; j < 16; j++) {
                    const uint64_t o = lo + j;
                    if (o < a || o >= b) continue;
                    const uint32_t byte = (w[j >> 2] >> (8 * (j & 3))) & 0xff;
                    st_u8(entries + o, byte);
                    for (int f = 0; f < N; f++)
                        if (S->peer_entries[f]) st_u8(S->peer_entries[f] + o, byte);
                }
            }
        }
        bar_sync(1, N_PRODUCER_THREADS);

        // ---- T6: bookkeeping + publish the tail (data before tail, invariant I1) --------
        if (gap) {
            // nothing is published after a gap tile: the next tile (at offset 0) carries it
            end = 0; hwm = L; pending_gap = true;
            bytes_rep += (b - a) * (uint64_t)(N - 1);
            if (tid == 0) { ctrl->hwm = hwm; ctrl->bytes_replicated = bytes_rep; }
            // the entry that wrapped stays first in the ring: re-run placement from offset 0
            // (descriptors are re-fetched; consumed is unchanged)
            continue;
        }
        uint64_t new_end = b;
        if (new_end == L) new_end = 0;                     // rule E1
        const uint32_t autoh = S->auto_head;
        tailpos = m ? a + S->rel[m - 1] : a;
        end = new_end;
        next_idx += m + autoh; consumed += m; published += m + autoh;
        auto_heads += autoh;
        prev_head = (autoh && m == 0);                     // never two HEAD entries in a row (dare_log.h:477-480)
        if (b > hwm) hwm = b;
        bytes_rep += (b - a) * (uint64_t)(N - 1);
        batches++;
        pending_gap = false;
        if (warp == 0) {
            if (lane < N && lane != me && cx->peer[lane]) {
                __threadfence_system();
                st_relaxed_sys(&reinterpret_cast<apus_loghdr_t *>(cx->peer[lane] + APUS_CTRL_BYTES)->end, new_end);
            }
            if (lane == 0) {
                hdr->end = new_end; hdr->tail = tailpos; hdr->old_end = new_end;
                ctrl->next_idx = next_idx; ctrl->consumed = consumed; ctrl->published = published;
                ctrl->hwm = hwm; ctrl->bytes_replicated = bytes_rep; ctrl->batches = batches;
                ctrl->auto_heads = auto_heads;
                const uint64_t h = S->pub_head;
                S->pub_cum[h & (PUB_RING - 1)] = published;
                S->pub_end[h & (PUB_RING - 1)] = new_end;
                S->pub_tickets[h & (PUB_RING - 1)] = consumed;
                S->pub_t0[h & (PUB_RING - 1)] = S->t_dequeue;
                __threadfence_block();
                S->pub_head = h + 1;
                S->published = published;
                st_relaxed_sys(&hw->consumed, consumed);
            }
        }
        last_progress = globaltimer_ns();
    }
    (void)pending_gap;
    if (tid == 0) { __threadfence_block(); S->producers_done = 1; }
}

// ---------------------------------------------------------------------------------
// FOLLOWER
// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t ring_dist(uint64_t from, uint64_t to, uint64_t L)
{
    return to >= from ? to - from : L - (from - to);
}

__device__ void follower_main(const apus_devctx_t *__restrict__ cx)
{
    FollowerShared *S = reinterpret_cast<FollowerShared *>(smem_raw);
    uint8_t *win = smem_raw + ((sizeof(FollowerShared) + 127u) & ~127u);
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int me = cx->idx, ldr = cx->leader_idx;
    apus_ctrl_t *ctrl = reinterpret_cast<apus_ctrl_t *>(cx->region);
    apus_loghdr_t *hdr = reinterpret_cast<apus_loghdr_t *>(cx->region + APUS_CTRL_BYTES);
    uint8_t *entries = cx->region + APUS_ENTRIES_OFF;
    apus_hostwords_t *hw = cx->hw;
    uint8_t *lregion = cx->peer[ldr];
    apus_ctrl_t *lctrl = reinterpret_cast<apus_ctrl_t *>(lregion);
    uint8_t *lentries = lregion + APUS_ENTRIES_OFF;
    const uint64_t L = cx->log_len;
    const bool fenced = (cx->flags & APUS_FLAG_FENCED_ACK) != 0;

    uint64_t old_end = hdr->old_end;     // walk position (dare_server.c:1795)
    uint64_t acked = ctrl->acked;
    uint64_t applied = hdr->apply;       // apply offset (== commit as far as I hold the entries)
    uint64_t pend_val = ctrl->pend_head_val, pend_end = ctrl->pend_head_end;
    uint64_t last_progress = globaltimer_ns();
    uint32_t spins = 0;

    for (;;) {
        if (tid == 0) {
            uint64_t e, c;
            uint32_t done = 0;
            for (;;) {
                e = ld_acquire_sys(&hdr->end);
                c = ld_relaxed_sys(&hdr->commit);
                const bool new_entries = (e != L) && (e != old_end);
                // commit moved, and I hold entries beyond what I applied
                const bool new_commit = (c != applied) && (old_end != L) && (applied != old_end);
                if (new_entries || new_commit) break;
                // bounded launch: the leader says how many entries exist in total
                if (cx->target != ~0ull && ld_acquire_sys(&ctrl->fin_target) == cx->target) {
                    const uint64_t fe = ld_relaxed_sys(&ctrl->fin_entries);
                    if (acked >= fe && (e == L || (applied == old_end && c == old_end))) { done = 1; break; }
                }
                if ((++spins & 0xffu) == 0) {
                    if (ld_relaxed_sys_u32(&hw->stop)) { done = 1; break; }
                    if (cx->target != ~0ull && globaltimer_ns() - last_progress > WATCHDOG_NS) {
                        st_relaxed_sys(&hw->error, APUS_KERR_WATCHDOG_FOLLOWER);
                        done = 1; break;
                    }
                }
            }
            S->end_seen = e; S->commit_seen = c; S->done = done;
        }
        __syncthreads();
        if (S->done) break;
        const uint64_t end_seen = S->end_seen, commit_seen = S->commit_seen;

        // ---- persist + ack every new entry in [old_end, end_seen) -----------------------
        while (end_seen != L && old_end != end_seen) {
            // window: contiguous bytes from old_end up to end_seen or the end of the ring
            if (tid == 0) {
                uint64_t lo = old_end;
                if (L - lo < APUS_HDR_BYTES) lo = 0;                 // log_get_entry: header does not fit -> 0
                uint64_t hi = (end_seen > lo) ? end_seen : L;       // wrapped: first run to the ring's end
                if (lo == end_seen) hi = lo;
                if (hi - (lo & ~15ull) > APUS_IMG_BYTES) hi = (lo & ~15ull) + APUS_IMG_BYTES;
                S->win_lo = lo; S->win_hi = hi;
            }
            __syncthreads();
            const uint64_t lo = S->win_lo, hi = S->win_hi;
            if (lo == hi) { old_end = lo; break; }
            const uint64_t lo16 = lo & ~15ull;
            const uint32_t nch = (uint32_t)(((hi + 15ull) & ~15ull) - lo16) >> 4;
            for (uint32_t c = tid; c < nch; c += nthr)
                reinterpret_cast<uint4 *>(win)[c] = ld_relaxed_sys_v4(entries + lo16 + 16ull * c);
            __syncthreads();
            // serial walk over headers in shared memory (log_get_entry / log_fit_entry / log_entry_len)
            if (tid == 0) {
                uint64_t off = lo;
                uint32_t n = 0;
                uint64_t next = off;
                bool wrapped = false;
                uint64_t hv = 0, he = L;
                while (off < hi) {
                    if (L - off < APUS_HDR_BYTES) { next = 0; wrapped = true; break; }   // jump to 0
                    if (hi - off < APUS_HDR_BYTES) { next = off; break; }                // header not in window yet
                    const uint8_t *e = win + (off - lo16);
                    const uint32_t ty = e[E_TYPE];
                    const uint32_t ln = (uint32_t)e[E_DATA] | ((uint32_t)e[E_DATA + 1] << 8);
                    const uint32_t es = entry_stride(ty, ln);
                    if (L - off < es) { next = 0; wrapped = true; break; }              // ghost: entry continues at 0
                    if (off + es > hi) { next = off; break; }                            // entry crosses the window
                    if (ty == T_HEAD) {                                                  // poll_config_entries (dare_server.c:2163-2170)
                        hv = 0;
                        for (int q = 7; q >= 0; q--) hv = (hv << 8) | e[E_DATA + q];
                        he = (off + es == L) ? 0 : off + es;
                    }
                    S->off[n++] = (uint32_t)(off - lo);
                    off += es;
                    next = off;
                }
                if (!wrapped && next == L) next = 0;      // rule E1 on walker offsets
                S->n = n; S->next = next;
                S->head_val = hv; S->head_end = he;
            }
            __syncthreads();
            const uint32_t n = S->n;
            // reply[me] = 1 in my copy and in the leader's copy (dare_ibv_rc.c:1833-1854)
            for (uint32_t k = tid; k < n; k += nthr) {
                const uint64_t at = lo + S->off[k] + E_REPLY + (uint64_t)me;
                st_relaxed_sys_u8(entries + at, 1);
                st_relaxed_sys_u8(lentries + at, 1);
            }
            __syncthreads();
            const uint64_t next = S->next;
            if (n == 0 && next == old_end) {
                // no progress possible inside this window: protocol error (entry larger than the window)
                if (tid == 0) st_relaxed_sys(&hw->error, APUS_KERR_BAD_ENTRY);
                old_end = end_seen;
                break;
            }
            if (S->head_end != L) { pend_val = S->head_val; pend_end = S->head_end; }
            acked += n;
            old_end = next;
            if (tid == 0) {
                if (fenced) __threadfence_system();     // reply bytes before the ack word
                st_relaxed_sys(&lctrl->ack[me], acked); // the word the leader's quorum ballot polls
                hdr->old_end = old_end;
                ctrl->acked = acked;
                ctrl->pend_head_val = pend_val; ctrl->pend_head_end = pend_end;
            }
            last_progress = globaltimer_ns();
        }

        // ---- follow the commit offset (invariant I4: never beyond what I hold) ----------
        if (old_end != L && applied != old_end && commit_seen != applied) {
            const uint64_t from = (applied == L) ? 0 : applied;
            const uint64_t held = ring_dist(from, old_end, L);          // entries I hold beyond `applied`
            uint64_t want = ring_dist(from, commit_seen, L);
            uint64_t to = commit_seen;
            if (want > held) { want = held; to = old_end; }             // the leader clamps the same way (dare_ibv_rc.c:1783-1787)
            if (want) {
                if (pend_end != L && ring_dist(from, pend_end, L) <= want && ring_dist(from, pend_end, L) > 0) {
                    // the HEAD entry is committed: adopt the head it carries (dare_server.c:2166-2169, 2182-2186)
                    if (tid == 0) { hdr->head = pend_val; ctrl->pend_head_end = L; }
                    pend_end = L;
                }
                applied = to;
                if (tid == 0) {
                    hdr->apply = applied;               // host-side apply (do_action) drains behind this
                    st_relaxed_sys(&lctrl->apply_off[me], applied);
                }
                last_progress = globaltimer_ns();
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------
extern "C" __global__ void __launch_bounds__(APUS_KERNEL_THREADS, 1)
apus_replica_kernel(const apus_role_t *__restrict__ roles)
{
    const apus_role_t r = roles[blockIdx.x];
    if (r.kind == APUS_ROLE_LEADER) leader_main(r.ctx);
    else if (r.kind == APUS_ROLE_FOLLOWER) follower_main(r.ctx);
}

extern "C" size_t apus_kernel_smem_bytes(void)
{
    size_t a = ((sizeof(LeaderShared) + 127u) & ~127u) + APUS_IMG_BYTES + 16;
    size_t b = ((sizeof(FollowerShared) + 127u) & ~127u) + APUS_IMG_BYTES + 16;
    return a > b ? a : b;
}

extern "C" cudaError_t apus_launch_roles(const apus_role_t *d_roles, int n_roles, cudaStream_t stream)
{
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    const size_t smem = apus_kernel_smem_bytes();
    if (dev < 64 && !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(apus_replica_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_set[dev] = true;
    }
    apus_replica_kernel<<<n_roles, APUS_KERNEL_THREADS, smem, stream>>>(d_roles);
    return cudaGetLastError();
}
