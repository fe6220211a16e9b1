/*
 * apus_gpu.h -- C ABI of the B200-native Paxos log-replication engine.
 *
 * This is the drop-in boundary for the ONE hot path of hku-systems/apus that this
 * repository accelerates: leader append -> replicate to every
 * follower -> majority ack -> commit.  It replaces the "lower" callee set the
 * reference's consensus thread calls into,
 *     src/include/dare/dare_ibv.h:141-201  (dare_ib_poll_tailq,
 *     dare_ib_write_remote_logs, dare_ib_send_entries_reply, ...) implemented by
 *     src/dare/dare_ibv_rc.c (RC queue pairs, RDMA WRITE/READ) and
 *     src/dare/dare_ibv_ud.c:780-790 (get_tailq_message),
 * together with the log placement of src/include/dare/dare_log.h.
 * Plain C types only: pointers, sizes, integers.  No CUDA or torch types.
 *
 * Each replica of a Paxos group is one `apus_replica_t`, bound to one GPU; its
 * consensus log (reference layout, byte for byte) and its ack / tail / commit
 * words live in that GPU's HBM.  The leader's hot loop and the followers' ack
 * loops are persistent sm_100a kernels (apus_b200/csrc/apus_kernels.cu); peers
 * are reached with P2P stores over NVLink (same process: peer access; one
 * process per replica: CUDA IPC handles exchanged with apus_replica_export /
 * apus_replica_connect -- the analogue of the raddr/rkey exchange in RC_SYN,
 * src/dare/dare_ibv_ud.c:1116-1119).
 *
 * The reference-facing "engine entry" symbols that src/proxy/proxy.c links
 * against (dare_server_init, is_leader, get_node_id, the tailhead queue) are
 * declared in apus_dare_entry.h and implemented on top of this ABI.
 *
 * Error convention follows the reference transport (dare_ibv_rc.c:27-29):
 * 0 = success, 1 = error, -1 = "retry later"; apus_last_error() gives the text.
 * There is NO CPU fallback: every entry point fails with 1 when no CUDA device
 * is usable.
 */
#ifndef APUS_GPU_H
#define APUS_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define APUS_ABI_VERSION 2

#define APUS_OK       0
#define APUS_ERROR    1
#define APUS_RETRY  (-1)

#define APUS_MAX_SERVER_COUNT 13            /* dare.h:26 */
#define APUS_LOG_SIZE (16384ull * 4096ull)  /* dare_log.h:76 LOG_SIZE */
#define APUS_ENTRY_HDR 64u                  /* sizeof(dare_log_entry_t) */

/* entry types: dare_log.h:21-24 and proxy.h:9-11 */
#define APUS_NOOP    0
#define APUS_CSM     1
#define APUS_CONFIG  2
#define APUS_HEAD    3
#define APUS_CONNECT 4
#define APUS_SEND    5
#define APUS_CLOSE   6

/* where the leader's submission ring lives */
#define APUS_RING_HOST_MAPPED 0   /* pinned host memory, read by the kernel over PCIe */
#define APUS_RING_DEVICE      1   /* HBM; the host fills it with cudaMemcpyAsync batches */

/* apus_config_t.flags */
#define APUS_F_FENCED_ACK   0x1u  /* follower: reply bytes visible before the ack word; off = ack as soon as
                                     the tail publish is observed, reply bytes follow (default off) */
#define APUS_F_DEVICE_STATS 0x2u  /* leader: record per-batch device-side commit latency */
#define APUS_F_AUTOPRUNE    0x4u  /* leader: device-side log pruning (force_log_pruning rule,
                                     dare_server.c:2069-2122): when the ring is a quarter full the
                                     kernel appends a HEAD entry carrying min(apply offsets) */
#define APUS_F_FOLLOWER_WALK 0x8u /* follower: find entry boundaries by walking the byte stream
                                     (log_get_entry/log_entry_len, as the reference follower does)
                                     instead of reading the leader-written offset index */
#define APUS_F_HOST_APPLY   0x10u /* follower: the apply offset reported to the leader's pruning rule is the one the
                                     HOST has replayed (apus_set_applied), as apply_committed_entries advances
                                     log->apply only after do_action (dare_server.c:1939-1962); off = apply follows
                                     commit on the device (nothing replays the log on the host) */
#define APUS_F_NO_EXPRESS   0x20u /* leader: no single-warp express path, every publish is fenced (A/B switch) */
#define APUS_F_FABRIC      0x100u /* the replica's HBM region is a VMM allocation that apus_group_multicast() can bind to an
                                     NVSwitch multicast object (replicas of one process, one GPU each) */
#define APUS_F_PROFILE      0x40u /* fine-grained device timestamps in the latency path (diagnostic runs only: each costs ~90 ns) */
#define APUS_F_EXPLICIT     0x80000000u /* flags are exactly as given (no defaults OR-ed in) */

typedef struct apus_replica apus_replica_t;

typedef struct apus_config {
    uint32_t struct_size;      /* sizeof(apus_config_t), for ABI growth */
    int32_t  device;           /* CUDA device ordinal hosting this replica */
    uint8_t  server_idx;       /* env server_idx (proxy.c:33-36) */
    uint8_t  group_size;       /* env group_size (proxy.c:37-40); 1..13 */
    uint8_t  leader_idx;       /* who leads in `term` (static until the control plane lands) */
    uint8_t  ring_mode;        /* APUS_RING_* (leader only) */
    uint32_t flags;            /* APUS_F_* */
    uint64_t term;             /* SID term stamped into entries (dare_server.h:53-61) */
    uint64_t log_size;         /* bytes of entries[]; 0 -> APUS_LOG_SIZE (reference) */
    uint32_t ring_slots;       /* submission descriptors, power of two; 0 -> default */
    uint32_t ring_bytes;       /* payload ring bytes, multiple of 4096; 0 -> default */
    uint32_t leader_ctas;      /* leader: worker CTAs (SMs) building tiles in parallel; 0 -> default */
    uint32_t reserved;
    /* ---- ABI 2 (struct_size tells which fields exist) ---- */
    uint32_t hb_period_us;     /* leader: heartbeat period, microseconds (hb_period of the config file,
                                  dare_server.c:763-791); 0 = no heartbeats */
    uint32_t hb_timeout_us;    /* follower: silence after which the leader is suspected (hb_timeout); 0 = never */
} apus_config_t;
#define APUS_CONFIG_SIZE_V1 48u

/* Opaque blob a replica publishes so that peers can map its HBM region.
 * Same process: carries the pointer; other process: a cudaIpcMemHandle_t. */
typedef struct apus_peer_handle {
    uint8_t bytes[128];
} apus_peer_handle_t;

/* the offsets of dare_log_t (dare_log.h:77-103) as this replica holds them */
typedef struct apus_log_offsets {
    uint64_t head, apply, commit, end, tail, old_end, old_commit, len;
} apus_log_offsets_t;

typedef struct apus_stats {
    uint64_t tickets_submitted;   /* requests accepted by apus_submit* */
    uint64_t tickets_consumed;    /* appended to the leader log by the kernel */
    uint64_t tickets_committed;   /* committed (majority acked), in log order */
    uint64_t entries_acked;       /* follower: entries acked to the leader */
    uint64_t bytes_replicated;    /* leader: sum over followers of entry bytes stored */
    uint64_t batches;             /* leader: replicate steps (tail publishes) */
    uint64_t kernel_launches;     /* launches of apus kernels that included this replica */
    uint64_t lat_samples;         /* device-side latency samples available */
    uint64_t auto_heads;          /* leader: HEAD entries appended by the device-side pruning rule */
    uint64_t entries_published;   /* leader: entries appended (tickets + auto HEAD entries) */
    uint64_t phase_ns[8];         /* leader profiling: ns waiting for requests, in T1..T6, and tile count */
    uint64_t turn_ns[8];          /* worker 0: [0..2] ns waiting for the claim lock / place turn / publish turn,
                                     [3] fast placements, [4] slow placements, [7] ns holding the place turn */
} apus_stats_t;

/* ---- library ------------------------------------------------------------------- */
int         apus_abi_version(void);
const char *apus_last_error(void);
int         apus_device_count(void);
/* NUMA node of the host memory next to a GPU, -1 if unknown: run the submitting threads (and allocate) there */
int         apus_device_numa_node(int device);

/* ---- replica life cycle --------------------------------------------------------- */
/* Allocates the HBM region (log header + entries + ctrl words), zeroed like
 * log_new() (dare_log.h:120-137: end = tail = old_end = len). */
int  apus_replica_create(const apus_config_t *cfg, apus_replica_t **out);
void apus_replica_destroy(apus_replica_t *r);

/* replaces the RC_SYN/SYNACK exchange of raddr+rkey (dare_ibv_ud.c:1116-1119) */
int  apus_replica_export(apus_replica_t *r, apus_peer_handle_t *out);
int  apus_replica_connect(apus_replica_t *r, uint8_t peer_idx, const apus_peer_handle_t *peer);

/* Replicas of ONE group, hosted by this process on pairwise different GPUs and created with APUS_F_FABRIC: bind their
 * regions to an NVSwitch multicast object.  The leader's replicate step then issues one `multimem.st` per 16 B chunk
 * (the switch fans it out to every replica, the leader's own copy included) instead of one store per replica. */
int  apus_group_multicast(apus_replica_t **rs, int n);

/* Launch the persistent kernel(s) for `n` replicas that live on the SAME device
 * in ONE fused launch (one CTA group per replica role).  The kernels return when
 * the cumulative ticket target is reached -- leader: that many requests committed;
 * follower: that many entries acked and applied -- or when apus_replicas_stop()
 * is called.  target_tickets == UINT64_MAX runs until stopped (service mode).
 * Asynchronous: returns after the launch. */
int  apus_replicas_launch(apus_replica_t **rs, int n, uint64_t target_tickets);
/* Wait for the launch that included `r` to finish; timeout_ms < 0 waits forever.
 * APUS_RETRY on timeout. */
int  apus_replica_wait(apus_replica_t *r, int64_t timeout_ms);
/* device time of the last finished launch that included r, in milliseconds (CUDA events) */
int  apus_replica_last_launch_ms(apus_replica_t *r, float *ms);
int  apus_replicas_stop(apus_replica_t **rs, int n);

/* ---- leader admission: the fields of tailq_entry_t (message.h:11-17) ------------- */
/* One request; `cmd` has `len` bytes (sm_cmd_t.cmd).  For APUS_CONFIG pass the
 * 16-byte dare_cid_t, for APUS_HEAD the 8-byte head offset, for APUS_NOOP nothing.
 * *ticket (optional) receives the 1-based position in the leader's append order.
 * APUS_RETRY when the submission ring is full. */
int  apus_submit(apus_replica_t *leader, uint8_t type, uint16_t connection_id, uint64_t req_id,
                 const void *cmd, uint16_t len, uint64_t *ticket);
/* n requests; payload k is payloads + k*stride (len[k] bytes).  conn/req arrays of n. */
int  apus_submit_batch(apus_replica_t *leader, uint32_t n, const uint8_t *types,
                       const uint16_t *connection_ids, const uint64_t *req_ids,
                       const uint16_t *lens, const void *payloads, size_t stride,
                       uint64_t *first_ticket);
/* n requests of ONE shape (type, connection, len; req_id = first_req_id + k; payload k at payloads + k*stride):
 * the bulk form of proxy.c:108-161's enqueue, filled by several host threads (env apus_submit_threads, default 8). */
int  apus_submit_uniform(apus_replica_t *leader, uint32_t n, uint8_t type, uint16_t connection_id,
                         uint64_t first_req_id, uint16_t len, const void *payloads, size_t stride,
                         uint64_t *first_ticket);
/* Device-generated requests (APUS_RING_DEVICE only): a fill kernel writes n SEND-like requests straight into the
 * HBM submission ring -- payload byte k of request req_id is apus_synth_byte(seed, req_id, k) -- so that a
 * benchmark can have its whole input resident in HBM without a host copy. */
int  apus_submit_synth(apus_replica_t *leader, uint32_t n, uint8_t type, uint16_t connection_id,
                       uint64_t first_req_id, uint16_t len, uint32_t seed, uint64_t *first_ticket);
uint8_t apus_synth_byte(uint32_t seed, uint64_t req_id, uint32_t k);
/* make everything submitted so far visible to the kernel (doorbell); apus_submit*
 * ring it themselves unless the replica was put in deferred mode */
int  apus_submit_defer(apus_replica_t *leader, int defer);
int  apus_submit_flush(apus_replica_t *leader);
/* ring the doorbell only up to `ticket` (<= submitted): requests beyond it stay resident but unseen */
int  apus_submit_release(apus_replica_t *leader, uint64_t ticket);

/* ---- commit observation (what update_state / do_action hang off) ---------------- */
uint64_t apus_committed_tickets(apus_replica_t *leader);
/* No CUDA call, just the pinned words the kernels keep up to date.  Leader: *offset = commit
 * offset, *count = tickets committed.  Follower: *offset = apply offset (everything before it
 * is committed and held by this replica -- what apply_committed_entries walks,
 * dare_server.c:1815-1974), *count = entries acked. */
int  apus_progress(apus_replica_t *r, uint64_t *offset, uint64_t *count);
/* spin until ticket is committed; APUS_RETRY on timeout */
int  apus_wait_committed(apus_replica_t *leader, uint64_t ticket, int64_t timeout_us);

/* n requests of payload_len bytes, ONE in flight at a time: submit, spin until committed
 * (what a proxy thread does, proxy.c:108-161); lat_ns[i] = host-clock nanoseconds of request i */
int  apus_closed_loop(apus_replica_t *leader, uint32_t n, uint16_t payload_len, uint16_t connection_id,
                      uint64_t first_req_id, uint32_t *lat_ns);

/* ---- inspection (parity tests, snapshots) --------------------------------------- */
int  apus_log_offsets(apus_replica_t *r, apus_log_offsets_t *out);
/* copy entries[off, off+len) of this replica's log image to host memory */
int  apus_log_read(apus_replica_t *r, uint64_t off, uint64_t len, void *dst);
int  apus_get_stats(apus_replica_t *r, apus_stats_t *out);
/* device-side commit latencies (ns), newest `max` samples; returns count in *n */
int  apus_latency_samples(apus_replica_t *r, uint32_t *dst_ns, uint32_t max, uint32_t *n);

/* follower, APUS_F_HOST_APPLY: the application has replayed the log up to `offset` (do_action done) */
int  apus_set_applied(apus_replica_t *r, uint64_t offset);
/* copy the committed-and-held range [from, to) of the circular log (it may wrap) into dst (capacity cap);
 * *got = bytes copied.  One or two device->host copies through a pinned buffer. */
int  apus_log_read_range(apus_replica_t *r, uint64_t from, uint64_t to, void *dst, uint64_t cap, uint64_t *got);
/* failure detector: 0 while the leader's heartbeats arrive, else 1 + the term whose leader fell silent */
uint64_t apus_leader_suspect(apus_replica_t *follower);
/* %globaltimer (ns) of the leader kernel's latest commit (device clock; step timing of resident kernels) */
uint64_t apus_last_commit_ns(apus_replica_t *leader);

/* ---- control plane on NVLink words (election, votes, log adjustment) ---------------------
 * Transport only: WHO votes for whom and when is decided by the caller (libapus_dare.so restates dare_server.c's
 * start_election / poll_vote_requests / poll_vote_count on top of these).  The words live in every replica's HBM
 * region next to its ack / tail slots (apus_layout.h: apus_ctlwords_t, the needed part of ctrl_data_t,
 * dare_server.h:121-138); a peer's words are written with a host-initiated copy through the same mapping the
 * kernels store through.  Replaces dare_ib_send_vote_request / replicate_vote / send_vote_ack
 * (dare_ibv_rc.c:969-1170), log_adjustment (dare_ibv_rc.c:1292-1451) and recover_log. */
typedef struct apus_ctl_view {
    uint64_t sid, leader_sid, adj_end, adj_count;
    uint64_t vote_ack[APUS_MAX_SERVER_COUNT];          /* commit offsets granted to me (log size = no vote) */
    struct { uint64_t sid, index, term, cid[2]; } vote_req[APUS_MAX_SERVER_COUNT];
} apus_ctl_view_t;
int  apus_ctl_read(apus_replica_t *r, apus_ctl_view_t *out);
int  apus_ctl_set_sid(apus_replica_t *r, uint64_t sid);
int  apus_ctl_reset_votes(apus_replica_t *r);                                  /* start_election: vote_ack[] := none */
int  apus_ctl_clear_vote_request(apus_replica_t *r, uint8_t from_idx);
int  apus_ctl_send_vote_request(apus_replica_t *r, uint8_t peer_idx, uint64_t sid, uint64_t index, uint64_t term,
                                const void *cid16);
int  apus_ctl_send_vote_ack(apus_replica_t *r, uint8_t candidate_idx, uint64_t commit);
/* follower: the heartbeat word the leader's kernel last wrote into this replica's region (term << 48 | beat counter;
 * dare_ibv_rc.c:868-958 writes the leader's SID into ctrl_data.hb[]).  The leader writes it whether or not this replica's
 * kernel runs, so a host that has stopped its kernel on a suspicion can tell a dead leader (the word stands still) from a
 * false positive (it moves) -- the reference's "false possitive => increase recomputed timeout", dare_server.c:781-796. */
int  apus_ctl_heartbeat(apus_replica_t *r, uint64_t *word);
/* idx and term of the last entry this replica holds (0,0 when the log is empty), its commit and end offsets; the
 * replica's kernel must be stopped (exclusive access, as dare_ib_revoke_log_access gives the reference) */
int  apus_ctl_last_entry(apus_replica_t *r, uint64_t *idx, uint64_t *term, uint64_t *commit, uint64_t *end);
/* elected leader, kernels stopped: bring follower `peer_idx` to my log -- find the last entry we share from its
 * commit offset on (log_find_remote_end_offset, dare_log.h:362-394), copy everything behind it (entry bytes and
 * offset index) peer to peer, and tell it to follow `sid` from there.  *resent = bytes copied. */
int  apus_ctl_adjust_follower(apus_replica_t *leader, uint8_t peer_idx, uint64_t sid, uint64_t *resent);
/* role and term for the next launch.  Becoming leader takes over the log as this replica holds it (entry counters,
 * tail, submission ring); becoming follower adopts what the new leader's adjustment left (apus_ctl_view.adj_*). */
int  apus_replica_set_role(apus_replica_t *r, uint8_t leader_idx, uint64_t term);
/* leader: liveness counters of the followers (their kernels bump them while polling); a counter that stands still
 * is a follower that is gone (HB replies, dare_ibv_rc.c:912-958 -> fail_count -> check_failure_count) */
int  apus_follower_beats(apus_replica_t *leader, uint64_t out[APUS_MAX_SERVER_COUNT]);
/* stop storing into a peer that is gone (dare_ib_disconnect_server, dare_server.c:1200) */
int  apus_replica_disconnect(apus_replica_t *r, uint8_t peer_idx);

/* control plane hooks used by pruning (log_pruning, dare_server.c:1996-2067).  apus_set_head writes the header word a
 * LAUNCH starts from; while kernels are resident the head moves the way the reference moves it: with the HEAD entry that
 * carries the new offset (submit APUS_HEAD with the 8-byte offset; the leader adopts it when it places the entry). */
int  apus_set_head(apus_replica_t *r, uint64_t head);
int  apus_remote_apply_offsets(apus_replica_t *leader, uint64_t out[APUS_MAX_SERVER_COUNT]);

#ifdef __cplusplus
}
#endif
#endif /* APUS_GPU_H */
