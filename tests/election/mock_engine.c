/*
 * tests/election/mock_engine.c -- TEST DOUBLE of the control-plane part of include/apus_gpu.h.
 * The election policy (apus_b200/csrc/dare_entry.c: elect -- the restatement of dare_server.c's start_election /
 * poll_vote_requests / poll_vote_count) only needs "words I can read and words I can write into a peer"; here those
 * words live in files mapped by every replica PROCESS instead of in HBM, so that the policy runs on a box without a GPU
 * (world size > 1, one process per replica).  Nothing of this is linked into the product.
 */
#define _GNU_SOURCE
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <unistd.h>

#include "apus_gpu.h"

typedef struct {
    uint64_t sid, leader_sid, adj_end, adj_count;
    uint64_t vote_ack[16];
    struct { uint64_t sid, index, term, cid[2], pad[3]; } vote_req[APUS_MAX_SERVER_COUNT];
    /* the replica's log, as far as an election looks at it */
    uint64_t last_idx, last_term, commit, end;
    /* what happened to it */
    uint64_t role_leader, role_term, launches, adjusted_by_plus1, disconnected_mask;
} blk_t;

struct apus_replica { int idx, n; blk_t *blk[APUS_MAX_SERVER_COUNT]; int in_flight; };
static char g_err[256];
const char *apus_last_error(void) { return g_err; }

apus_replica_t *mock_open(const char *dir, int idx, int n, uint64_t last_idx, uint64_t last_term, uint64_t commit, uint64_t end, uint64_t sid)
{
    apus_replica_t *r = calloc(1, sizeof *r);
    r->idx = idx; r->n = n;
    for (int i = 0; i < n; i++) {
        char path[512];
        snprintf(path, sizeof path, "%s/ctl%d.bin", dir, i);
        int fd = open(path, O_RDWR | O_CREAT, 0666);
        if (fd < 0 || ftruncate(fd, sizeof(blk_t)) != 0) { perror(path); exit(2); }
        r->blk[i] = mmap(NULL, sizeof(blk_t), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        close(fd);
    }
    blk_t *b = r->blk[idx];
    b->sid = sid; b->last_idx = last_idx; b->last_term = last_term; b->commit = commit; b->end = end;
    for (int i = 0; i < 16; i++) b->vote_ack[i] = APUS_LOG_SIZE;
    __sync_synchronize();
    r->in_flight = 1;
    return r;
}

int apus_replicas_stop(apus_replica_t **rs, int n) { for (int i = 0; i < n; i++) rs[i]->in_flight = 0; return APUS_OK; }
int apus_replicas_launch(apus_replica_t **rs, int n, uint64_t t) { (void)t; for (int i = 0; i < n; i++) { rs[i]->in_flight = 1; rs[i]->blk[rs[i]->idx]->launches++; } return APUS_OK; }

int apus_ctl_read(apus_replica_t *r, apus_ctl_view_t *out)
{
    blk_t *b = r->blk[r->idx];
    __sync_synchronize();
    memset(out, 0, sizeof *out);
    out->sid = b->sid; out->leader_sid = b->leader_sid; out->adj_end = b->adj_end; out->adj_count = b->adj_count;
    for (int i = 0; i < APUS_MAX_SERVER_COUNT; i++) {
        out->vote_ack[i] = b->vote_ack[i];
        out->vote_req[i].index = b->vote_req[i].index; out->vote_req[i].term = b->vote_req[i].term;
        out->vote_req[i].cid[0] = b->vote_req[i].cid[0]; out->vote_req[i].cid[1] = b->vote_req[i].cid[1];
        out->vote_req[i].sid = b->vote_req[i].sid;
    }
    return APUS_OK;
}
int apus_ctl_set_sid(apus_replica_t *r, uint64_t sid) { r->blk[r->idx]->sid = sid; __sync_synchronize(); return APUS_OK; }
int apus_ctl_reset_votes(apus_replica_t *r) { for (int i = 0; i < 16; i++) r->blk[r->idx]->vote_ack[i] = APUS_LOG_SIZE; __sync_synchronize(); return APUS_OK; }
int apus_ctl_clear_vote_request(apus_replica_t *r, uint8_t from) { r->blk[r->idx]->vote_req[from].sid = 0; __sync_synchronize(); return APUS_OK; }
int apus_ctl_send_vote_request(apus_replica_t *r, uint8_t peer, uint64_t sid, uint64_t index, uint64_t term, const void *cid16)
{
    blk_t *p = r->blk[peer];
    p->vote_req[r->idx].index = index; p->vote_req[r->idx].term = term;
    if (cid16) memcpy(p->vote_req[r->idx].cid, cid16, 16);
    __sync_synchronize();
    p->vote_req[r->idx].sid = sid;                 /* the word the voter polls goes last */
    __sync_synchronize();
    return APUS_OK;
}
int apus_ctl_send_vote_ack(apus_replica_t *r, uint8_t cand, uint64_t commit) { r->blk[cand]->vote_ack[r->idx] = commit; __sync_synchronize(); return APUS_OK; }
int apus_ctl_last_entry(apus_replica_t *r, uint64_t *idx, uint64_t *term, uint64_t *commit, uint64_t *end)
{
    if (r->in_flight) { snprintf(g_err, sizeof g_err, "the replica's kernel must be stopped"); return APUS_ERROR; }
    blk_t *b = r->blk[r->idx];
    *idx = b->last_idx; *term = b->last_term;
    if (commit) *commit = b->commit;
    if (end) *end = b->end;
    return APUS_OK;
}
int apus_ctl_adjust_follower(apus_replica_t *r, uint8_t peer, uint64_t sid, uint64_t *resent)
{
    if (r->in_flight) { snprintf(g_err, sizeof g_err, "the leader's kernel must be stopped"); return APUS_ERROR; }
    blk_t *me = r->blk[r->idx], *p = r->blk[peer];
    if (resent) *resent = me->end > p->end ? me->end - p->end : 0;
    p->last_idx = me->last_idx; p->last_term = me->last_term; p->end = me->end;      /* the follower holds my log now */
    p->adjusted_by_plus1 = (uint64_t)r->idx + 1;
    p->adj_end = me->end; p->adj_count = me->last_idx;
    __sync_synchronize();
    p->leader_sid = sid;
    __sync_synchronize();
    return APUS_OK;
}
int apus_replica_set_role(apus_replica_t *r, uint8_t leader, uint64_t term)
{
    /* MOCK_TAKEOVER_DELAY_US: a winner that is slow to take over (descheduled, a long role change): its voters' patience
     * runs out and they stand in a higher term before it announces itself */
    const char *d = getenv("MOCK_TAKEOVER_DELAY_US");
    if (d && leader == r->idx) usleep((useconds_t)atol(d));
    r->blk[r->idx]->role_leader = leader; r->blk[r->idx]->role_term = term; return APUS_OK;
}
int apus_replica_disconnect(apus_replica_t *r, uint8_t peer) { r->blk[r->idx]->disconnected_mask |= 1ull << peer; return APUS_OK; }
/* MOCK_LEADER_ALIVE=<term>: the "dead" leader still beats (a false positive of the failure detector) */
int apus_ctl_heartbeat(apus_replica_t *r, uint64_t *word)
{
    static uint64_t beat;
    const char *t = getenv("MOCK_LEADER_ALIVE");
    (void)r;
    *word = t ? ((uint64_t)atol(t) << 48) | ++beat : 77;
    return APUS_OK;
}

/* ---- the rest of the ABI dare_entry.c references: never reached by the election harness ---- */
#define STUB(sig) sig { snprintf(g_err, sizeof g_err, "mock: not part of the election harness"); return APUS_ERROR; }
int apus_device_count(void) { return 0; }
STUB(int apus_replica_create(const apus_config_t *c, apus_replica_t **o))
void apus_replica_destroy(apus_replica_t *r) { (void)r; }
STUB(int apus_replica_export(apus_replica_t *r, apus_peer_handle_t *o))
int apus_replica_connect(apus_replica_t *r, uint8_t p, const apus_peer_handle_t *h) { (void)r; (void)p; (void)h; return APUS_OK; }
STUB(int apus_submit(apus_replica_t *l, uint8_t t, uint16_t c, uint64_t q, const void *m, uint16_t n, uint64_t *k))
int apus_submit_defer(apus_replica_t *l, int d) { (void)l; (void)d; return APUS_OK; }
int apus_submit_flush(apus_replica_t *l) { (void)l; return APUS_OK; }
uint64_t apus_committed_tickets(apus_replica_t *l) { (void)l; return 0; }
STUB(int apus_progress(apus_replica_t *r, uint64_t *o, uint64_t *c))
STUB(int apus_log_read_range(apus_replica_t *r, uint64_t f, uint64_t t, void *d, uint64_t c, uint64_t *g))
int apus_set_applied(apus_replica_t *r, uint64_t o) { (void)r; (void)o; return APUS_OK; }
uint64_t apus_leader_suspect(apus_replica_t *r) { (void)r; return 0; }
STUB(int apus_follower_beats(apus_replica_t *l, uint64_t o[APUS_MAX_SERVER_COUNT]))
