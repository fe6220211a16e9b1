/*
 * tests/hostlogic/pump_harness.c -- runs the PRODUCT's host pumps (apus_b200/csrc/dare_entry.c: follower_pump and
 * leader_pump, included here as source) on a box without a GPU.  The engine behind them is the test double below:
 *
 *   follower  the replica's circular log is a sequence of STAGES written by the test (the oracle's image of the ring and
 *             its commit offset after every few requests); `apus_progress` reports the stage's commit offset,
 *             `apus_log_read_range` copies out of the staged ring exactly like the real call (one or two pieces, capped),
 *             and the next stage is installed only when the pump has reported the current one applied
 *             (`apus_set_applied`) -- the protocol by which APUS_F_HOST_APPLY keeps the leader from lapping the host.
 *             Every store_cmd / do_action callback is written down; the test compares them with the request stream.
 *   leader    application threads enqueue tailq entries under tailq_lock the way proxy.c:108-161 does and spin until
 *             "committed"; `apus_submit` writes down what it was given and hands out tickets, `apus_committed_tickets`
 *             trails the submissions; store_cmd / update_state callbacks are written down.
 *
 *   pump_harness follower <dir> <log_len> <n_stages> <read_cap>      reads <dir>/stage<k>.bin, <dir>/stage<k>.commit
 *   pump_harness leader   <dir> <n_threads> <n_requests_per_thread> <payload_len>
 * Output: <dir>/calls.txt.  Nothing of this is linked into the product.
 */
#include "../../apus_b200/csrc/dare_entry.c"

struct apus_replica { int unused; };
static struct apus_replica g_mock;
static char g_merr[256];
const char *apus_last_error(void) { return g_merr; }
static FILE *g_calls;

/* ---- follower side ---- */
static uint8_t *g_ring;
static uint64_t g_L, g_commit, g_applied, g_read_cap;
static int g_stage = -1, g_nstages;
static const char *g_dir;
static unsigned g_reads, g_two_piece_reads, g_capped_reads;

static int load_stage(int k)
{
    char path[600];
    snprintf(path, sizeof path, "%s/stage%d.bin", g_dir, k);
    FILE *f = fopen(path, "rb");
    if (!f || fread(g_ring, 1, g_L, f) != g_L) { fprintf(stderr, "cannot read %s\n", path); exit(2); }
    fclose(f);
    snprintf(path, sizeof path, "%s/stage%d.commit", g_dir, k);
    f = fopen(path, "r");
    unsigned long long c = 0;
    if (!f || fscanf(f, "%llu", &c) != 1) { fprintf(stderr, "cannot read %s\n", path); exit(2); }
    fclose(f);
    g_commit = c; g_stage = k;
    return 0;
}

int apus_progress(apus_replica_t *r, uint64_t *off, uint64_t *cnt)
{
    (void)r;
    if (g_stage < 0 || g_applied == g_commit) {
        if (g_stage + 1 < g_nstages) load_stage(g_stage + 1);
        else g_terminate = 1;                                  /* everything replayed: let the pump return */
    }
    *off = g_commit; *cnt = 0;
    return APUS_OK;
}

int apus_log_read_range(apus_replica_t *r, uint64_t from, uint64_t to, void *dst, uint64_t cap, uint64_t *got)
{
    (void)r;
    if (from >= g_L || to >= g_L) { snprintf(g_merr, sizeof g_merr, "range beyond the log"); return APUS_ERROR; }
    if (g_read_cap && cap > g_read_cap) cap = g_read_cap;      /* the test shrinks the buffer to exercise "cut by the buffer" */
    uint64_t n1 = to >= from ? to - from : g_L - from, n2 = to >= from ? 0 : to;
    if (n1 > cap) { n1 = cap; n2 = 0; g_capped_reads++; }
    if (n1 + n2 > cap) { n2 = cap - n1; g_capped_reads++; }
    memcpy(dst, g_ring + from, n1);
    if (n2) { memcpy((uint8_t *)dst + n1, g_ring, n2); g_two_piece_reads++; }
    *got = n1 + n2;
    g_reads++;
    return APUS_OK;
}

int apus_set_applied(apus_replica_t *r, uint64_t o)
{
    (void)r;
    if (o >= g_L) { snprintf(g_merr, sizeof g_merr, "offset beyond the log"); return APUS_ERROR; }
    g_applied = o;
    fprintf(g_calls, "P %llu\n", (unsigned long long)o);
    return APUS_OK;
}
uint64_t apus_leader_suspect(apus_replica_t *r) { (void)r; return 0; }

static uint64_t fnv(const uint8_t *p, size_t n) { uint64_t h = 1469598103934665603ull; for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ull; } return h; }
static void rec_store(void *data, void *arg)
{
    (void)arg;
    const uint8_t *d = data;                                   /* the entry from clt_id on (dare_server.c:1802) */
    uint16_t clt, len; memcpy(&clt, d, 2); memcpy(&len, d + 24, 2);
    fprintf(g_calls, "S %u %u %u\n", (unsigned)clt, (unsigned)d[2], (unsigned)len);
}
static void rec_action(uint16_t clt, uint8_t type, size_t n, void *data, void *arg)
{
    (void)arg;
    fprintf(g_calls, "A %u %u %zu %016llx\n", (unsigned)clt, (unsigned)type, n, (unsigned long long)fnv(data, n));
}

/* ---- leader side ---- */
static pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER;
static uint64_t g_tickets, g_flushed;
static volatile uint64_t g_done_state;                         /* update_state calls so far == the proxy's highest_rec */

int apus_submit(apus_replica_t *l, uint8_t t, uint16_t c, uint64_t q, const void *m, uint16_t n, uint64_t *k)
{
    (void)l;
    pthread_mutex_lock(&g_mu);
    *k = ++g_tickets;
    fprintf(g_calls, "T %llu %u %u %llu %u %016llx\n", (unsigned long long)*k, (unsigned)t, (unsigned)c, (unsigned long long)q,
            (unsigned)n, (unsigned long long)fnv(m ? m : (const void *)"", m ? n : 0));
    pthread_mutex_unlock(&g_mu);
    return APUS_OK;
}
int apus_submit_defer(apus_replica_t *l, int d) { (void)l; (void)d; return APUS_OK; }
int apus_submit_flush(apus_replica_t *l) { (void)l; pthread_mutex_lock(&g_mu); g_flushed = g_tickets; pthread_mutex_unlock(&g_mu); return APUS_OK; }
/* commits trail the doorbell: everything flushed is committed, one call later */
uint64_t apus_committed_tickets(apus_replica_t *l)
{
    (void)l;
    static uint64_t last;
    pthread_mutex_lock(&g_mu);
    uint64_t c = last; last = g_flushed;
    pthread_mutex_unlock(&g_mu);
    return c;
}
static void rec_update(void *arg) { (void)arg; g_done_state++; }
static void rec_store_leader(void *data, void *arg)
{
    (void)arg;
    const uint8_t *d = data;
    uint16_t clt, len; memcpy(&clt, d, 2); memcpy(&len, d + 24, 2);
    pthread_mutex_lock(&g_mu);
    fprintf(g_calls, "S %u %u %u\n", (unsigned)clt, (unsigned)d[2], (unsigned)len);
    pthread_mutex_unlock(&g_mu);
}

static void *pump_thread(void *a) { (void)a; leader_pump(0); return NULL; }

typedef struct { int id, nreq, plen; } app_t;
static void *app_thread(void *a)
{
    /* proxy.c:108-161: build the tailq entry, append it under tailq_lock, remember my record number, spin until the
     * state machine has caught up with it */
    app_t *ap = a;
    uint64_t req = 0;
    for (int i = 0; i < ap->nreq + 2; i++) {
        tailq_entry_t *e = calloc(1, sizeof *e);
        e->type = i == 0 ? APUS_CONNECT : (i == ap->nreq + 1 ? APUS_CLOSE : APUS_SEND);
        e->connection_id = (uint16_t)ap->id;
        e->req_id = ++req;
        if (e->type == APUS_SEND) {
            e->cmd.len = (uint16_t)ap->plen;
            for (int k = 0; k < ap->plen; k++) e->cmd.cmd[k] = (uint8_t)(ap->id * 131 + i * 31 + k);
        }
        pthread_spin_lock(&tailq_lock);
        static uint64_t cur_rec;                               /* (under the lock, like proxy->cur_rec) */
        const uint64_t mine = ++cur_rec;
        TAILQ_INSERT_TAIL(&tailhead, e, entries);
        pthread_spin_unlock(&tailq_lock);
        while (g_done_state < mine) sched_yield();
    }
    return NULL;
}

/* ---- the rest of the ABI dare_entry.c references: not reached by the pumps ---- */
#define STUB(sig) sig { snprintf(g_merr, sizeof g_merr, "mock: not part of the pump harness"); return APUS_ERROR; }
int apus_device_count(void) { return 0; }
STUB(int apus_replica_create(const apus_config_t *c, apus_replica_t **o))
void apus_replica_destroy(apus_replica_t *r) { (void)r; }
STUB(int apus_replica_export(apus_replica_t *r, apus_peer_handle_t *o))
STUB(int apus_replica_connect(apus_replica_t *r, uint8_t p, const apus_peer_handle_t *h))
int apus_replicas_stop(apus_replica_t **rs, int n) { (void)rs; (void)n; return APUS_OK; }
int apus_replicas_launch(apus_replica_t **rs, int n, uint64_t t) { (void)rs; (void)n; (void)t; return APUS_OK; }
STUB(int apus_follower_beats(apus_replica_t *l, uint64_t o[APUS_MAX_SERVER_COUNT]))
STUB(int apus_ctl_read(apus_replica_t *r, apus_ctl_view_t *o))
STUB(int apus_ctl_set_sid(apus_replica_t *r, uint64_t s))
STUB(int apus_ctl_reset_votes(apus_replica_t *r))
STUB(int apus_ctl_clear_vote_request(apus_replica_t *r, uint8_t f))
STUB(int apus_ctl_send_vote_request(apus_replica_t *r, uint8_t p, uint64_t s, uint64_t i, uint64_t t, const void *c))
STUB(int apus_ctl_send_vote_ack(apus_replica_t *r, uint8_t c, uint64_t k))
STUB(int apus_ctl_last_entry(apus_replica_t *r, uint64_t *a, uint64_t *b, uint64_t *c, uint64_t *d))
STUB(int apus_ctl_adjust_follower(apus_replica_t *l, uint8_t f, uint64_t s, uint64_t *b))
STUB(int apus_replica_set_role(apus_replica_t *r, uint8_t l, uint64_t t))
STUB(int apus_replica_disconnect(apus_replica_t *r, uint8_t p))

int main(int argc, char **argv)
{
    if (argc < 6) return 2;
    g_dir = argv[2];
    char path[600];
    snprintf(path, sizeof path, "%s/calls.txt", g_dir);
    g_calls = fopen(path, "w");
    if (!g_calls) return 2;
    g_log = stdout;
    g_rep = &g_mock;
    g_tk_type = calloc(TK_RING, 1);
    if (!strcmp(argv[1], "follower")) {
        g_L = strtoull(argv[3], NULL, 0); g_nstages = atoi(argv[4]); g_read_cap = strtoull(argv[5], NULL, 0);
        g_ring = malloc(g_L);
        g_log_len = g_L; g_n = 3; g_idx = 1; g_leader_idx = 0;
        g_in.store_cmd = rec_store; g_in.do_action = rec_action;
        int rc = follower_pump(g_L);
        fprintf(g_calls, "END rc=%d apply=%llu next_idx=%llu reads=%u two_piece=%u capped=%u\n", rc, (unsigned long long)g_apply,
                (unsigned long long)g_apply_next_idx, g_reads, g_two_piece_reads, g_capped_reads);
    } else {
        int nthr = atoi(argv[3]), nreq = atoi(argv[4]), plen = atoi(argv[5]);
        g_n = 1; g_idx = 0; g_leader_idx = 0; g_live_mask = 1;      /* (a group of one appends no CONFIG prologue) */
        g_in.store_cmd = rec_store_leader; g_in.update_state = rec_update;
        pthread_spin_init(&tailq_lock, PTHREAD_PROCESS_PRIVATE);
        TAILQ_INIT(&tailhead);
        pthread_t th[64]; app_t ap[64];
        for (int i = 0; i < nthr; i++) { ap[i].id = i; ap[i].nreq = nreq; ap[i].plen = plen; pthread_create(&th[i], NULL, app_thread, &ap[i]); }
        pthread_t pump;
        pthread_create(&pump, NULL, pump_thread, NULL);
        for (int i = 0; i < nthr; i++) pthread_join(th[i], NULL);
        g_terminate = 1;
        pthread_join(pump, NULL);
        fprintf(g_calls, "END tickets=%llu update_state=%llu\n", (unsigned long long)g_tickets, (unsigned long long)g_done_state);
    }
    fclose(g_calls);
    return 0;
}
