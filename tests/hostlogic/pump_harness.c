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
 *   membership  a leader of three whose follower p2 stops beating: the failure detector must remove it (disconnect, CONFIG
 *             entry without it); then a replacement process (mode `joiner`, a second process sharing the rendezvous
 *             directory) asks to join: the leader snapshots the state machine through the proxy callbacks, maps the
 *             joiner, sends its log, appends the CONFIG entry that puts the slot back and answers; the joiner loads the
 *             snapshot and knows where to follow from (leader_check_followers / leader_serve_join / join_group).
 *
 *   pump_harness follower <dir> <log_len> <n_stages> <read_cap>      reads <dir>/stage<k>.bin, <dir>/stage<k>.commit
 *   pump_harness leader   <dir> <n_threads> <n_requests_per_thread> <payload_len>
 *   pump_harness membership <dir> <p2_dies_after_ms> 0 0
 *   pump_harness joiner     <dir> 0 0 0
 *   pump_harness config     <dir> <libconfig file> 0 0        prints what read_dare_config took from the file
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
    if (t == APUS_CONFIG && m) {                              /* dare_cid_t image: size at byte 8, bitmask at bytes 12..15 */
        uint32_t mask; memcpy(&mask, (const uint8_t *)m + 12, 4);
        fprintf(g_calls, "CONFIG %llu size=%u mask=%u\n", (unsigned long long)*k, (unsigned)((const uint8_t *)m)[8], mask);
    } else
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

/* ---- membership: what the failure detector and the join service ask of the engine ---- */
static uint64_t g_t0_us, g_p2_dies_us, g_beat[APUS_MAX_SERVER_COUNT];
static unsigned g_stops, g_launches;
static volatile unsigned g_disconnected_mask;
int apus_replicas_stop(apus_replica_t **rs, int n) { (void)rs; (void)n; g_stops++; return APUS_OK; }
int apus_replicas_launch(apus_replica_t **rs, int n, uint64_t t) { (void)rs; (void)n; (void)t; g_launches++; return APUS_OK; }
int apus_follower_beats(apus_replica_t *l, uint64_t o[APUS_MAX_SERVER_COUNT])
{
    (void)l;
    g_beat[1]++;                                               /* p1 is alive */
    if (now_us() - g_t0_us < g_p2_dies_us) g_beat[2]++;        /* p2's kernel stops polling when its process dies */
    memcpy(o, g_beat, sizeof g_beat);
    return APUS_OK;
}
int apus_replica_disconnect(apus_replica_t *r, uint8_t p)
{
    (void)r;
    pthread_mutex_lock(&g_mu); fprintf(g_calls, "D %u\n", (unsigned)p); pthread_mutex_unlock(&g_mu);
    g_disconnected_mask |= 1u << p;
    return APUS_OK;
}
int apus_replica_connect(apus_replica_t *r, uint8_t p, const apus_peer_handle_t *h)
{
    (void)r;
    pthread_mutex_lock(&g_mu); fprintf(g_calls, "C %u %016llx\n", (unsigned)p, (unsigned long long)fnv((const uint8_t *)h, sizeof *h)); pthread_mutex_unlock(&g_mu);
    return APUS_OK;
}
int apus_ctl_last_entry(apus_replica_t *r, uint64_t *idx, uint64_t *term, uint64_t *commit, uint64_t *end)
{
    (void)r; *idx = 41; *term = 5; *commit = 2624; *end = 2624;
    return APUS_OK;
}
int apus_ctl_adjust_follower(apus_replica_t *l, uint8_t f, uint64_t sid, uint64_t *resent)
{
    (void)l; *resent = 2624;
    pthread_mutex_lock(&g_mu); fprintf(g_calls, "J %u %llu\n", (unsigned)f, (unsigned long long)sid); pthread_mutex_unlock(&g_mu);
    return APUS_OK;
}
static uint32_t snap_size(void *arg) { (void)arg; return 1000; }
static void snap_fill(void *buf, void *arg) { (void)arg; for (int i = 0; i < 1000; i++) ((uint8_t *)buf)[i] = (uint8_t)(i * 7 + 3); }
static int snap_apply(void *buf, uint32_t n, void *arg)
{
    (void)arg;
    fprintf(g_calls, "SNAP %u %016llx\n", n, (unsigned long long)fnv(buf, n));
    return 0;
}

/* ---- the rest of the ABI dare_entry.c references: not reached here ---- */
#define STUB(sig) sig { snprintf(g_merr, sizeof g_merr, "mock: not part of the pump harness"); return APUS_ERROR; }
int apus_device_count(void) { return 0; }
STUB(int apus_replica_create(const apus_config_t *c, apus_replica_t **o))
void apus_replica_destroy(apus_replica_t *r) { (void)r; }
STUB(int apus_replica_export(apus_replica_t *r, apus_peer_handle_t *o))
STUB(int apus_ctl_read(apus_replica_t *r, apus_ctl_view_t *o))
STUB(int apus_ctl_set_sid(apus_replica_t *r, uint64_t s))
STUB(int apus_ctl_reset_votes(apus_replica_t *r))
STUB(int apus_ctl_clear_vote_request(apus_replica_t *r, uint8_t f))
STUB(int apus_ctl_send_vote_request(apus_replica_t *r, uint8_t p, uint64_t s, uint64_t i, uint64_t t, const void *c))
STUB(int apus_ctl_send_vote_ack(apus_replica_t *r, uint8_t c, uint64_t k))
STUB(int apus_replica_set_role(apus_replica_t *r, uint8_t l, uint64_t t))
STUB(int apus_ctl_heartbeat(apus_replica_t *r, uint64_t *w))

int main(int argc, char **argv)
{
    if (argc < 6) return 2;
    g_dir = argv[2];
    char path[600];
    snprintf(path, sizeof path, "%s/%s", g_dir, !strcmp(argv[1], "joiner") ? "joiner_calls.txt" : "calls.txt");
    g_calls = fopen(path, "w");
    if (!g_calls) return 2;
    g_log = stdout;
    g_rep = &g_mock;
    g_tk_type = calloc(TK_RING, 1);
    if (!strcmp(argv[1], "config")) {
        read_dare_config(argv[3]);
        printf("hb_period=%g elec_timeout_low=%llu elec_timeout_high=%llu\n", cfg_hb_period, (unsigned long long)cfg_elec_low,
               (unsigned long long)cfg_elec_high);
    } else if (!strcmp(argv[1], "follower")) {
        g_L = strtoull(argv[3], NULL, 0); g_nstages = atoi(argv[4]); g_read_cap = strtoull(argv[5], NULL, 0);
        g_ring = malloc(g_L);
        g_log_len = g_L; g_n = 3; g_idx = 1; g_leader_idx = 0;
        g_in.store_cmd = rec_store; g_in.do_action = rec_action;
        int rc = follower_pump(g_L);
        fprintf(g_calls, "END rc=%d apply=%llu next_idx=%llu reads=%u two_piece=%u capped=%u\n", rc, (unsigned long long)g_apply,
                (unsigned long long)g_apply_next_idx, g_reads, g_two_piece_reads, g_capped_reads);
    } else if (!strcmp(argv[1], "membership")) {
        g_n = 3; g_idx = 0; g_leader_idx = 0; g_live_mask = 7; g_term = 5;
        cfg_hb_period = 0.002;                                      /* hb_timeout() = 10 periods, floored at 20 ms */
        snprintf(g_env_rdv, sizeof g_env_rdv, "%s/rdv", g_dir);
        mkdir(g_env_rdv, 0777);
        g_in.get_db_size = snap_size; g_in.create_db_snapshot = snap_fill; g_in.update_state = rec_update; g_in.store_cmd = rec_store_leader;
        g_t0_us = now_us(); g_p2_dies_us = strtoull(argv[3], NULL, 0) * 1000ull;
        pthread_spin_init(&tailq_lock, PTHREAD_PROCESS_PRIVATE);
        TAILQ_INIT(&tailhead);
        pthread_t pump;
        pthread_create(&pump, NULL, pump_thread, NULL);
        char mark[700];
        for (int k = 0; k < 10000 && !(g_disconnected_mask & 4u); k++) usleep(1000);
        snprintf(mark, sizeof mark, "%s/removed", g_dir);
        FILE *f = fopen(mark, "w"); if (f) { fprintf(f, "%llu\n", (unsigned long long)(now_us() - g_t0_us)); fclose(f); }
        snprintf(mark, sizeof mark, "%s/join2.ack", g_env_rdv);
        for (int k = 0; k < 20000 && access(mark, F_OK); k++) usleep(1000);
        usleep(100000);
        g_terminate = 1;
        pthread_join(pump, NULL);
        pthread_mutex_lock(&g_mu);
        fprintf(g_calls, "END stops=%u launches=%u live_mask=%u removed_mask=%u\n", g_stops, g_launches, g_live_mask, g_removed_mask);
        pthread_mutex_unlock(&g_mu);
    } else if (!strcmp(argv[1], "joiner")) {
        g_n = 3; g_idx = 2; g_leader_idx = 2; g_L = 1ull << 30;
        snprintf(g_env_rdv, sizeof g_env_rdv, "%s/rdv", g_dir);
        g_in.apply_db_snapshot = snap_apply;
        char path[700];
        apus_peer_handle_t h;
        memset(&h, 0xA7, sizeof h);                                 /* (dare_server_init publishes the handle before join_group) */
        snprintf(path, sizeof path, "%s/r2.handle", g_env_rdv);
        FILE *f = fopen(path, "wb"); if (f) { fwrite(&h, sizeof h, 1, f); fclose(f); }
        int rc = join_group();
        fprintf(g_calls, "END rc=%d leader=%u term=%llu apply=%llu next_idx=%llu live_mask=%u handle=%016llx\n", rc, (unsigned)g_leader_idx,
                (unsigned long long)g_term, (unsigned long long)g_apply, (unsigned long long)g_apply_next_idx, g_live_mask,
                (unsigned long long)fnv((const uint8_t *)&h, sizeof h));
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
