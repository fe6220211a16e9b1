/*
 * dare_entry.c -- libapus_dare.so: the engine-entry symbols APUS's proxy.c links against
 * (include/apus_dare_entry.h), implemented on the C ABI of include/apus_gpu.h.
 *
 * Replaces, from the caller's point of view, the reference's libdare.a entry points
 * (src/dare/dare_server.c:173-241 dare_server_init, :243-255 dare_server_shutdown,
 * :2299-2307 is_leader/get_node_id) and the TAILQ globals of message.h:20-22.  The thread
 * that runs dare_server_init is "the DARE thread": it is the only one that touches the
 * engine and the only one that invokes proxy callbacks, as in the reference.
 */
#define _GNU_SOURCE
#include <dirent.h>
#include <errno.h>
#include <execinfo.h>
#include <signal.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <sys/time.h>
#include <time.h>
#include <unistd.h>

#include "apus_dare_entry.h"
#include "apus_gpu.h"

/* the one real definition of the reference's header-defined globals (message.h:20-22) */
struct apus_tailhead_t tailhead;
pthread_spinlock_t tailq_lock;
/* dare_log.h:27 declares it extern and every includer of that header (proxy.c among them)
 * references it; the reference defines it in dare_server.c */
int prev_log_entry_head;

static FILE *g_log;
static apus_replica_t *g_rep;
static volatile int g_leader, g_started, g_terminate;
static uint8_t g_idx, g_n, g_leader_idx;
static uint64_t g_term = 1;                 /* term of a clean first election */
static uint32_t g_live_mask;                /* servers of the configuration (dare_cid_t.bitmask) */
static uint32_t g_removed_mask;             /* servers a new leader found dead: removed with a CONFIG entry */
static uint64_t g_log_len = APUS_LOG_SIZE;
static uint64_t g_apply, g_apply_next_idx;  /* follower apply walk: survives a change of leader */
static unsigned g_false_positives;          /* suspicions of a leader that turned out to be alive (elect) */
static dare_server_input_t g_in;

/* dare_global_config of the libconfig file (config-dare.c:12-52; target/nodes.local.cfg) */
static double   cfg_hb_period = 0.01;             /* seconds */
static uint64_t cfg_elec_low = 100000, cfg_elec_high = 300000;   /* microseconds */

/* SID = [TERM|L|IDX] (dare_server.h:46-61) */
#define SID_IDX(s)   ((uint8_t)((s) & 0xFF))
#define SID_L(s)     (((s) >> 8) & 1)
#define SID_TERM(s)  ((s) >> 9)
#define SID_MAKE(term, l, idx) ((((uint64_t)(term)) << 9) | ((uint64_t)((l) ? 1 : 0) << 8) | (uint64_t)(idx))

#define LOGT(fmt, ...) do { struct timeval _tv; gettimeofday(&_tv, NULL); \
    fprintf(g_log, "[%lu:%06lu] " fmt, (unsigned long)_tv.tv_sec, (unsigned long)_tv.tv_usec, ##__VA_ARGS__); fflush(g_log); } while (0)

/* test / tooling hook: the replica behind this process (inspect it through include/apus_gpu.h) */
void *apus_dare_replica(void) { return g_rep; }
uint64_t apus_dare_term(void) { return g_term; }

int is_leader(void) { return g_started && g_leader; }          /* dare_server.c:2299-2302 */
uint8_t get_node_id(void) { return g_idx; }                    /* dare_server.c:2304-2307 */

static void int_handler(int sig) { (void)sig; g_terminate = 1; }   /* dare_server.c:2309-2315 */

void dare_server_shutdown(void)
{
    if (g_rep) {
        apus_replica_t *rs[1] = { g_rep };
        apus_replicas_stop(rs, 1);
        apus_replica_destroy(g_rep);
        g_rep = NULL;
    }
    g_started = 0;
    if (g_log && g_log != stdout) fclose(g_log);
    pthread_exit(NULL);
}

/* ---- peer handle rendezvous: the stand-in for RC_SYN/SYNACK/ACK (dare_ibv_ud.c:1168-1416) ---- */
static int rendezvous(const char *dir, const apus_peer_handle_t *mine, apus_peer_handle_t *all)
{
    char path[512], tmp[560];
    mkdir(dir, 0777);
    snprintf(path, sizeof path, "%s/r%u.handle", dir, (unsigned)g_idx);
    snprintf(tmp, sizeof tmp, "%s.tmp.%d", path, (int)getpid());
    FILE *f = fopen(tmp, "wb");
    if (!f) return 1;
    fwrite(mine, sizeof *mine, 1, f);
    fclose(f);
    if (rename(tmp, path)) return 1;
    all[g_idx] = *mine;
    for (unsigned i = 0; i < g_n; i++) {
        if (i == g_idx) continue;
        snprintf(path, sizeof path, "%s/r%u.handle", dir, i);
        for (int tries = 0;; tries++) {
            f = fopen(path, "rb");
            if (f) {
                size_t got = fread(&all[i], sizeof all[i], 1, f);
                fclose(f);
                if (got == 1) break;
            }
            if (g_terminate || tries > 60000) return 1;      /* 60 s */
            usleep(1000);
        }
    }
    return 0;
}

/* apus_segv_trace=1: print a backtrace on SIGSEGV/SIGBUS/SIGABRT before dying (the host application may
 * install its own handler later; this one covers the engine's start-up inside an LD_PRELOADed process) */
static void segv_trace(int sig)
{
    void *bt[64];
    int n = backtrace(bt, 64);
    static const char msg[] = "apus: fatal signal in an engine-hosting process, backtrace:\n";
    if (write(2, msg, sizeof msg - 1) < 0) { }
    backtrace_symbols_fd(bt, n, 2);
    signal(sig, SIG_DFL);
    raise(sig);
}
/* The engine's own environment variables are read ONCE, when the library is loaded (before the application's
 * main()): the DARE thread must not call getenv() -- the hosting application may rewrite environ concurrently
 * (redis-server's setproctitle does: clearenv + setenv at the top of main), and getenv() racing with that
 * crashes.  The reference reads its variables in proxy.c:33-58 on the main thread for the same reason. */
static int g_env_gpu = -1, g_env_leader = 0;
static int g_env_colocate;                 /* 1: this (leader) process also hosts the followers' replicas: kernels only */
static long g_env_hb_us = -1, g_env_hbto_us = -1, g_env_elec_lo = -1, g_env_elec_hi = -1;
static unsigned long long g_env_log_size;
static char g_env_rdv[256];
__attribute__((constructor)) static void engine_env_init(void)
{
    const char *s;
    if ((s = getenv("apus_gpu"))) g_env_gpu = atoi(s);
    if ((s = getenv("apus_leader"))) g_env_leader = atoi(s);
    if ((s = getenv("apus_log_size"))) g_env_log_size = strtoull(s, NULL, 0);
    if ((s = getenv("apus_colocate_followers"))) g_env_colocate = atoi(s);
    if ((s = getenv("apus_hb_period_us"))) g_env_hb_us = atol(s);
    if ((s = getenv("apus_hb_timeout_us"))) g_env_hbto_us = atol(s);
    if ((s = getenv("apus_elec_timeout_us"))) { g_env_elec_lo = atol(s); const char *c = strchr(s, ','); g_env_elec_hi = c ? atol(c + 1) : 2 * g_env_elec_lo; }
    if ((s = getenv("apus_rendezvous"))) snprintf(g_env_rdv, sizeof g_env_rdv, "%s", s);
    else snprintf(g_env_rdv, sizeof g_env_rdv, "/tmp/apus-rdv-%u", (unsigned)getuid());
    if (getenv("apus_segv_trace")) { signal(SIGSEGV, segv_trace); signal(SIGBUS, segv_trace); signal(SIGABRT, segv_trace); }
}

static void cid_image(uint8_t out[16], uint32_t bitmask);
static uint64_t now_us(void);
static int launch_self(void);
static int csm_like(uint8_t type) { return !(type == APUS_NOOP || type == APUS_CONFIG || type == APUS_HEAD); }

#define TK_RING (1u << 20)
static uint8_t *g_tk_type;        /* type of every ticket in flight, indexed by ticket & (TK_RING-1) */
static uint64_t g_join_ticket;    /* ticket of a CONFIG entry appended by leader_serve_join (the pump counts it as submitted) */

/* ---- join: a replaced server comes back into an EMPTY SLOT of the configuration ----------------
 * The reference's joiner multicasts a JOIN request, the leader adds it with a CONFIG entry, the joiner fetches a snapshot
 * of the state machine from a server (recover_sm: proxy get_db_size / create_db_snapshot there, apply_db_snapshot here,
 * dare_server.c:598-721) and then the log (recover_log), dare_ibv_ud.c:952-1087, dare_server.c:1883-1937.  On one box the
 * request / reply messages are files next to the peer handles (the rendezvous directory already stands in for the UD
 * bootstrap); the snapshot is the proxy's own (its stored records, replayed into the joining application), the log
 * comes over NVLink with the same peer-to-peer copy the log adjustment uses.  Growing the group (CID_EXTENDED ->
 * TRANSIT -> STABLE, size[1]) is not built: a join is accepted for a slot that a CONFIG entry has emptied. */
static int read_handle(unsigned i, apus_peer_handle_t *h)
{
    char path[512];
    snprintf(path, sizeof path, "%s/r%u.handle", g_env_rdv, i);
    FILE *f = fopen(path, "rb");
    if (!f) return 1;
    size_t got = fread(h, sizeof *h, 1, f);
    fclose(f);
    return got == 1 ? 0 : 1;
}

static uint64_t g_beat_val[APUS_MAX_SERVER_COUNT], g_beat_seen[APUS_MAX_SERVER_COUNT], g_last_beat_scan;
static uint64_t g_last_join_scan;
/* leader: serve one pending join request, if any.  Called between batches with nothing in flight. */
static void leader_serve_join(uint64_t submitted, uint64_t applied)
{
    const uint64_t now = now_us();
    if (now - g_last_join_scan < 20000) return;
    g_last_join_scan = now;
    for (unsigned i = 0; i < g_n; i++) {
        if (i == g_idx || (g_live_mask & (1u << i))) continue;
        char req[512], ack[560], snap[512];
        snprintf(req, sizeof req, "%s/join%u.req", g_env_rdv, i);
        if (access(req, F_OK)) continue;
        if (submitted != applied) return;                     /* quiesce first: everything appended is committed and applied */
        LOGT("JOIN request from p%u\n", i);
        /* the state machine snapshot (poll_sm_requests, dare_server.c:598-652) */
        uint32_t len = g_in.get_db_size ? g_in.get_db_size(g_in.up_para) : 0;
        uint8_t *buf = (uint8_t *)malloc(len ? len : 1);
        if (len && g_in.create_db_snapshot) g_in.create_db_snapshot(buf, g_in.up_para);
        snprintf(snap, sizeof snap, "%s/snap%u.bin", g_env_rdv, i);
        FILE *f = fopen(snap, "wb");
        if (f) { fwrite(buf, 1, len, f); fclose(f); }
        free(buf);
        LOGT("   # snapshot len = %u\n", len);
        /* the log: stop my kernel, map the joiner, copy my log behind its (empty) one, take it into the configuration */
        apus_replica_t *rs[1] = { g_rep };
        apus_replicas_stop(rs, 1);
        apus_peer_handle_t h;
        uint64_t resent = 0, last_idx = 0, last_term = 0, commit = 0, end = 0;
        int ok = read_handle(i, &h) == 0 && apus_replica_connect(g_rep, (uint8_t)i, &h) == APUS_OK &&
                 apus_ctl_last_entry(g_rep, &last_idx, &last_term, &commit, &end) == APUS_OK &&
                 apus_ctl_adjust_follower(g_rep, (uint8_t)i, SID_MAKE(g_term, 1, g_idx), &resent) == APUS_OK;
        if (!ok) LOGT("join of p%u failed: %s\n", i, apus_last_error());
        if (launch_self() != APUS_OK) { LOGT("launch: %s\n", apus_last_error()); g_terminate = 1; return; }
        unlink(req);
        if (!ok) return;
        g_live_mask |= 1u << i; g_removed_mask &= ~(1u << i);
        g_beat_seen[i] = 0; g_beat_val[i] = 0;                /* a new process: its liveness counter starts over */
        { uint64_t b[APUS_MAX_SERVER_COUNT]; if (apus_follower_beats(g_rep, b) == APUS_OK) g_beat_val[i] = b[i]; }
        uint8_t cid[16];
        cid_image(cid, g_live_mask);
        uint64_t t = 0;
        if (apus_submit(g_rep, APUS_CONFIG, 0, 0, cid, 0, &t) == APUS_OK) { g_tk_type[t & (TK_RING - 1)] = APUS_CONFIG; g_join_ticket = t; apus_submit_flush(g_rep); }
        snprintf(ack, sizeof ack, "%s/join%u.ack.tmp", g_env_rdv, i);
        f = fopen(ack, "w");
        if (f) {
            fprintf(f, "%u %llu %llu %llu %u %u\n", (unsigned)g_idx, (unsigned long long)g_term, (unsigned long long)commit,
                    (unsigned long long)(last_idx + 1), len, g_live_mask);
            fclose(f);
            char fin[512];
            snprintf(fin, sizeof fin, "%s/join%u.ack", g_env_rdv, i);
            rename(ack, fin);
        }
        LOGT("p%u joined: %llu log bytes sent, CONFIG entry appended (bitmask %03x)\n", i, (unsigned long long)resent, g_live_mask);
        return;
    }
}

/* joiner: announce myself, wait for the leader's reply, load the snapshot; returns 0 when ready to run as a follower */
static int join_group(void)
{
    char path[560], tmp[600];
    snprintf(path, sizeof path, "%s/join%u.ack", g_env_rdv, (unsigned)g_idx);
    unlink(path);
    snprintf(path, sizeof path, "%s/join%u.req", g_env_rdv, (unsigned)g_idx);
    snprintf(tmp, sizeof tmp, "%s.tmp", path);
    FILE *f = fopen(tmp, "w");
    if (!f) return 1;
    fprintf(f, "%d\n", (int)getpid());
    fclose(f);
    rename(tmp, path);
    LOGT("JOIN request sent (slot p%u)\n", (unsigned)g_idx);
    snprintf(path, sizeof path, "%s/join%u.ack", g_env_rdv, (unsigned)g_idx);
    unsigned lead = 0, snaplen = 0, mask = 0;
    unsigned long long term = 0, apply = 0, next_idx = 0;
    for (int tries = 0;; tries++) {
        f = fopen(path, "r");
        if (f) {
            int got = fscanf(f, "%u %llu %llu %llu %u %u", &lead, &term, &apply, &next_idx, &snaplen, &mask);
            fclose(f);
            if (got == 6) break;
        }
        if (g_terminate || tries > 60000) { LOGT("no reply to the JOIN request\n"); return 1; }
        usleep(1000);
    }
    /* recover_sm: the snapshot goes through the proxy into the application (apply_db_snapshot, proxy.c:300-339) */
    if (snaplen) {
        snprintf(path, sizeof path, "%s/snap%u.bin", g_env_rdv, (unsigned)g_idx);
        uint8_t *buf = (uint8_t *)malloc(snaplen);
        f = fopen(path, "rb");
        size_t got = f ? fread(buf, 1, snaplen, f) : 0;
        if (f) fclose(f);
        if (got != snaplen) { LOGT("snapshot short: %zu of %u bytes\n", got, snaplen); free(buf); return 1; }
        if (g_in.apply_db_snapshot && g_in.apply_db_snapshot(buf, snaplen, g_in.up_para)) { LOGT("apply_db_snapshot failed\n"); free(buf); return 1; }
        free(buf);
    }
    g_leader_idx = (uint8_t)lead; g_term = term; g_live_mask = mask;
    g_apply = apply; g_apply_next_idx = next_idx;
    apus_set_applied(g_rep, apply);
    LOGT("joined: leader p%u, term %llu, snapshot of %u bytes applied, log follows from offset %llu (idx %llu)\n", lead, term, snaplen, apply, next_idx);
    return 0;
}

/* leader: failure detector for FOLLOWERS.  A follower's kernel bumps its liveness counter in my HBM while it polls
 * (the HB replies of dare_ibv_rc.c:912-958); a counter that stands still for hb_timeout is a server that is gone:
 * it is disconnected and removed from the configuration with a CONFIG entry (check_failure_count,
 * dare_server.c:1189-1228), so that nothing stores into its memory any more and a replacement can join its slot. */
static void leader_check_followers(uint64_t *submitted)
{
    const uint64_t now = now_us();
    if (now - g_last_beat_scan < 5000) return;
    g_last_beat_scan = now;
    uint64_t b[APUS_MAX_SERVER_COUNT];
    if (apus_follower_beats(g_rep, b) != APUS_OK) return;
    uint64_t timeout = g_env_hbto_us > 0 ? (uint64_t)g_env_hbto_us : (uint64_t)(10.0 * cfg_hb_period * 1e6);
    if (timeout < 20000) timeout = 20000;
    for (unsigned i = 0; i < g_n; i++) {
        if (i == g_idx || !(g_live_mask & (1u << i))) continue;
        if (b[i] != g_beat_val[i]) { g_beat_val[i] = b[i]; g_beat_seen[i] = now; continue; }
        if (!g_beat_seen[i] || now - g_beat_seen[i] < timeout) continue;       /* never seen yet (still starting), or recent */
        LOGT("REMOVE SERVER p%u\n", i);                                          /* dare_server.c:1203 */
        apus_replica_t *rs[1] = { g_rep };
        apus_replicas_stop(rs, 1);
        apus_replica_disconnect(g_rep, (uint8_t)i);
        g_live_mask &= ~(1u << i); g_removed_mask |= 1u << i;
        if (launch_self() != APUS_OK) { LOGT("launch: %s\n", apus_last_error()); g_terminate = 1; return; }
        uint8_t cid[16];
        cid_image(cid, g_live_mask);
        uint64_t t = 0;
        if (apus_submit(g_rep, APUS_CONFIG, 0, 0, cid, 0, &t) == APUS_OK) { g_tk_type[t & (TK_RING - 1)] = APUS_CONFIG; *submitted = t; apus_submit_flush(g_rep); }
    }
}

/* ---- leader pump ---------------------------------------------------------------------------- */

static void leader_pump(int elected)
{
    uint64_t submitted = 0, applied = 0;
    uint8_t image[64];
    struct apus_tailhead_t pending;                 /* popped from the shared queue, not yet in the engine */
    TAILQ_INIT(&pending);
    if (g_n > 1) {
        /* the election winner's blank CONFIG entry (dare_server.c:1412-1421): it is what lets entries of earlier
         * terms commit, as a prefix of an entry of this term */
        uint8_t cid[16];
        cid_image(cid, g_live_mask);
        uint64_t t = 0;
        if (apus_submit(g_rep, APUS_CONFIG, 0, 0, cid, 0, &t) != APUS_OK) { LOGT("cannot append CONFIG: %s\n", apus_last_error()); return; }
        g_tk_type[t & (TK_RING - 1)] = APUS_CONFIG;
        submitted = t;
        if (g_removed_mask & g_live_mask) {
            /* servers that did not take part in the election are removed from the configuration
             * (check_failure_count, dare_server.c:1189-1228: CID_SERVER_RM + a CONFIG entry) */
            g_live_mask &= ~g_removed_mask;
            cid_image(cid, g_live_mask);
            if (apus_submit(g_rep, APUS_CONFIG, 0, 0, cid, 0, &t) != APUS_OK) { LOGT("cannot append CONFIG: %s\n", apus_last_error()); return; }
            g_tk_type[t & (TK_RING - 1)] = APUS_CONFIG;
            submitted = t;
        }
    }
    if (!elected) LOGT("[T%llu] LEADER\n", (unsigned long long)g_term);   /* benchmarks/run.sh:52 greps for "] LEADER" */
    g_leader = 1;
    apus_submit_defer(g_rep, 1);
    while (!g_terminate) {
        /* get_tailq_message (dare_ibv_ud.c:780-790): lock, pop FIFO, append, free -- the lock (a spinlock every
         * application thread takes in proxy.c:108-161) is held only to unhook what is queued: O(1), no engine call,
         * no BerkeleyDB put under it */
        if (!TAILQ_EMPTY(&tailhead)) {               /* (racy peek; the splice below is under the lock) */
            pthread_spin_lock(&tailq_lock);
            TAILQ_CONCAT(&pending, &tailhead, entries);
            pthread_spin_unlock(&tailq_lock);
        }
        /* pass 1: hand the requests to the engine in FIFO order and ring the doorbell ONCE */
        tailq_entry_t *batch[256];
        int nb = 0;
        while (nb < 256 && !TAILQ_EMPTY(&pending) && submitted - applied < TK_RING - 2) {
            tailq_entry_t *n3 = TAILQ_FIRST(&pending);
            uint64_t t = 0;
            int rc = apus_submit(g_rep, n3->type, n3->connection_id, n3->req_id, n3->cmd.cmd, n3->cmd.len, &t);
            if (rc == APUS_RETRY) break;                       /* ring full: flush, drain commits, retry */
            if (rc != APUS_OK) { LOGT("apus_submit: %s\n", apus_last_error()); g_terminate = 1; break; }
            g_tk_type[t & (TK_RING - 1)] = n3->type;
            submitted = t;
            TAILQ_REMOVE(&pending, n3, entries);
            batch[nb++] = n3;
        }
        if (nb) apus_submit_flush(g_rep);
        /* pass 2: while the GPUs replicate, persist.  persist_new_entries (dare_server.c:1802): store_cmd(&entry->clt_id):
         * the bytes of the entry from clt_id on, as they are when the leader persists (sender/reply unset) */
        for (int k = 0; k < nb; k++) {
            tailq_entry_t *n3 = batch[k];
            memset(image, 0, sizeof image);
            memcpy(image, &n3->connection_id, 2);
            image[2] = n3->type;
            memcpy(image + 24, &n3->cmd.len, 2);
            if (g_in.store_cmd) g_in.store_cmd(image, g_in.up_para);
            free(n3);
        }
        if (g_n > 1 && !g_env_colocate) leader_check_followers(&submitted);
        if (g_n > 1 && g_live_mask != ((1u << g_n) - 1u)) {
            leader_serve_join(submitted, applied);
            if (g_join_ticket) { submitted = g_join_ticket; g_join_ticket = 0; }
        }
        /* apply_committed_entries, leader branch (dare_server.c:1851-1861, 1951-1952) */
        uint64_t c = apus_committed_tickets(g_rep);
        while (applied < c) {
            applied++;
            if (csm_like(g_tk_type[applied & (TK_RING - 1)]) && g_in.update_state) g_in.update_state(g_in.up_para);
        }
    }
}

/* ---- follower pump: apply_committed_entries, follower branch (dare_server.c:1815-1967) -------- */
static int follower_pump(uint64_t L)
{
    uint64_t apply = g_apply, next_idx = g_apply_next_idx;
    const size_t cap = 1u << 20;
    uint8_t *buf = (uint8_t *)malloc(cap + 65536 + 64);
    uint32_t idle = 0;
    while (!g_terminate) {
        uint64_t off = 0, cnt = 0;
        if (apus_progress(g_rep, &off, &cnt) != APUS_OK) { LOGT("%s\n", apus_last_error()); break; }
        if (off == apply) {
            if (apus_leader_suspect(g_rep)) { g_apply = apply; g_apply_next_idx = next_idx; free(buf); return 1; }
            if (++idle > 2000) usleep(20);       /* spin first: the commit word is plain host memory */
            continue;
        }
        idle = 0;
        /* everything in [apply, off) is committed and held by this replica: ONE range read (two copies when the
         * range wraps) into a pinned buffer, then walk it with the reference's rules */
        uint64_t got = 0;
        if (apus_log_read_range(g_rep, apply, off, buf, cap + 65536 + 64, &got) != APUS_OK) { LOGT("%s\n", apus_last_error()); break; }
        uint64_t o = apply, p = 0;               /* log offset / buffer position */
        while (p < got && !g_terminate) {
            if (L - o < APUS_ENTRY_HDR) { p += L - o; o = 0; continue; }               /* log_get_entry: no room for a header */
            if (got - p < APUS_ENTRY_HDR) break;
            const uint8_t *e = buf + p;
            const uint8_t type = e[26];
            uint16_t len; memcpy(&len, e + 48, 2);
            const uint32_t stride = csm_like(type) ? 64u + len : 64u;                  /* log_entry_len */
            if (L - o < stride) { p += L - o; o = 0; continue; }                       /* ghost header: the entry is at 0 */
            if (got - p < stride) break;                                                 /* cut by the buffer: next round */
            uint64_t idx; memcpy(&idx, e, 8);
            if (next_idx && idx != next_idx) {
                /* the bytes are not the entry that should be here: the host fell a whole ring behind (cannot happen
                 * while the leader prunes by the replayed offset, APUS_F_HOST_APPLY) -- never replay garbage */
                LOGT("FATAL: apply walk expected idx %llu, found %llu at offset %llu\n", (unsigned long long)next_idx,
                     (unsigned long long)idx, (unsigned long long)o);
                g_terminate = 1;
                break;
            }
            next_idx = idx + 1;
            if (csm_like(type)) {
                uint16_t clt; memcpy(&clt, e + 24, 2);
                if (g_in.store_cmd) g_in.store_cmd((void *)(e + 24), g_in.up_para);
                if (g_in.do_action) g_in.do_action(clt, type, len, (void *)(e + 50), g_in.up_para);
                if (idx % 10000 == 0) { uint64_t term; memcpy(&term, e + 8, 8);
                    LOGT("APPLY LOG ENTRY: (%llu; %llu)\n", (unsigned long long)idx, (unsigned long long)term); }
            }
            p += stride; o += stride;
            if (o == L) o = 0;
        }
        if (o != apply) {
            apply = o;
            apus_set_applied(g_rep, apply);      /* log->apply moves only after do_action (dare_server.c:1939-1962) */
        }
    }
    g_apply = apply; g_apply_next_idx = next_idx;
    free(buf);
    return 0;
}

/* ---- dare_global_config: the three values of the config file this engine uses (config-dare.c:12-52).  The
 *      interposer links libconfig for proxy.c; the engine only needs `key = number;` inside that one group ---- */
static void read_dare_config(const char *path)
{
    /* libconfig syntax as far as this group needs it: `name = value;` or `name : value;` settings in any layout (several on
     * one line -- benchmarks/run_gpu.sh writes them so --, one per line like target/nodes.local.cfg), comments `#`, `//`
     * and C style, integers with an optional L suffix; only settings inside `dare_global_config = { ... }` count */
    FILE *f = path && path[0] ? fopen(path, "r") : NULL;
    if (!f) return;
    char *txt = (char *)malloc(65536);
    size_t n = txt ? fread(txt, 1, 65535, f) : 0;
    fclose(f);
    if (!txt) return;
    txt[n] = 0;
    for (char *p = txt; *p; p++) {                               /* blank the comments out */
        if (*p == '"') { for (p++; *p && *p != '"'; p++) if (*p == '\\' && p[1]) p++; if (!*p) break; continue; }
        if (*p == '#' || (*p == '/' && p[1] == '/')) { while (*p && *p != '\n') *p++ = ' '; if (!*p) break; continue; }
        if (*p == '/' && p[1] == '*') { while (*p && !(*p == '*' && p[1] == '/')) *p++ = ' '; if (!*p) break; *p++ = ' '; *p = ' '; }
    }
    char *g = strstr(txt, "dare_global_config");
    char *open = g ? strchr(g, '{') : NULL, *close = open ? strchr(open, '}') : NULL;
    if (open && close) {
        *close = 0;
        for (char *p = open + 1; *p; ) {
            while (*p && !((*p >= 'a' && *p <= 'z') || *p == '_')) p++;
            char key[64]; size_t k = 0;
            while ((*p >= 'a' && *p <= 'z') || (*p >= '0' && *p <= '9') || *p == '_') { if (k < sizeof key - 1) key[k++] = *p; p++; }
            key[k] = 0;
            while (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r') p++;
            if (*p != '=' && *p != ':') continue;
            char *end = NULL;
            const double v = strtod(p + 1, &end);
            if (end == p + 1) { p++; continue; }
            p = end;
            if (!strcmp(key, "hb_period")) cfg_hb_period = v;
            else if (!strcmp(key, "elec_timeout_low")) cfg_elec_low = (uint64_t)v;
            else if (!strcmp(key, "elec_timeout_high")) cfg_elec_high = (uint64_t)v;
        }
    }
    free(txt);
}

static uint64_t now_us(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (uint64_t)ts.tv_sec * 1000000ull + (uint64_t)ts.tv_nsec / 1000ull;
}

/* random_election_timeout (dare_server.c:1237-1250) */
static uint64_t random_election_timeout_us(void)
{
    struct timeval tv;
    gettimeofday(&tv, NULL);
    srand48((long)(g_idx + 1) * (long)((tv.tv_sec % 100) * 1000000 + tv.tv_usec));
    const uint64_t span = cfg_elec_high > cfg_elec_low ? cfg_elec_high - cfg_elec_low : 1;
    return (uint64_t)(lrand48() % span) + cfg_elec_low;
}

static void cid_image(uint8_t out[16], uint32_t bitmask)
{
    memset(out, 0, 16);
    out[8] = g_n;                                 /* dare_cid_t {epoch, size[2], state, bitmask}, dare_config.h */
    memcpy(out + 12, &bitmask, 4);
}

static int launch_self(void)
{
    apus_replica_t *rs[1] = { g_rep };
    return apus_replicas_launch(rs, 1, UINT64_MAX);
}

/* ---- leader failure: election, log adjustment, change of role ---------------------------------------
 * Restates, on the control words of include/apus_gpu.h, the reference's
 *   start_election      dare_server.c:1264-1320   term+1, SID [t|0|me], votes cleared, vote request (sid, last idx, last term)
 *   poll_vote_requests  dare_server.c:1524-1743   best SID wins, up-to-date test on (term, idx), one vote per term, vote ack = my commit
 *   poll_vote_count     dare_server.c:1322-1400   size/2+1 votes incl. my own -> [t|1|me], "[T<t>] LEADER"
 *   log_adjustment      dare_ibv_rc.c:1292-1451   (apus_ctl_adjust_follower) and the follower's side of it.
 * Returns 0 when this replica runs again in its new role (g_leader_idx / g_term updated), 1 to give up. */
static int elect(void)
{
    apus_replica_t *rs[1] = { g_rep };
    apus_replicas_stop(rs, 1);                              /* exclusive access to my log (dare_ib_revoke_log_access) */
    const uint8_t dead = g_leader_idx;
    LOGT("[T%llu] leader p%u fell silent: failure detector fired\n", (unsigned long long)g_term, (unsigned)dead);
    {   /* a false positive?  (hb_receive_cb, dare_server.c:781-796: a beat that arrives after the leader was declared failed
         * only lengthens the timeout.)  The leader's kernel writes its beat into my region whether or not my kernel runs;
         * its own kernel pauses for milliseconds around a change of the peer set (removal, join).  If the word moves
         * within a few periods, and is of the term I follow, the leader is alive: keep following. */
        uint64_t hb0 = 0, hb1 = 0;
        const uint64_t period_us = (uint64_t)(cfg_hb_period * 1e6);
        if (apus_ctl_heartbeat(g_rep, &hb0) == APUS_OK) {
            for (int k = 0; k < 2 && !g_terminate; k++) {               /* a live leader beats every period: two are enough */
                usleep((useconds_t)(period_us < 500 ? 500 : period_us));
                if (apus_ctl_heartbeat(g_rep, &hb1) != APUS_OK) break;
                if (hb1 != hb0 && (hb1 >> 48) == (g_term & 0xffffull)) {
                    g_false_positives++;
                    LOGT("false positive => p%u is alive (beat %llu -> %llu): keep following, %u so far\n", (unsigned)dead,
                         (unsigned long long)(hb0 & 0xffffffffffffull), (unsigned long long)(hb1 & 0xffffffffffffull), g_false_positives);
                    if (launch_self() != APUS_OK) { LOGT("launch: %s\n", apus_last_error()); return 1; }
                    return 0;
                }
            }
        }
    }
    apus_ctl_view_t v;
    if (apus_ctl_read(g_rep, &v) != APUS_OK) { LOGT("%s\n", apus_last_error()); return 1; }
    uint64_t sid = v.sid;
    sid &= ~(1ull << 8);                                    /* no leader known any more */
    uint64_t last_idx = 0, last_term = 0, commit = 0, end = 0;
    if (apus_ctl_last_entry(g_rep, &last_idx, &last_term, &commit, &end) != APUS_OK) { LOGT("%s\n", apus_last_error()); return 1; }
    uint8_t cid[16];
    cid_image(cid, g_live_mask);
    {   /* servers that joined after my start-up published their handles later: map them now (no-op for known peers) */
        apus_peer_handle_t h;
        for (unsigned i = 0; i < g_n; i++)
            if (i != g_idx && i != dead && read_handle(i, &h) == 0) apus_replica_connect(g_rep, (uint8_t)i, &h);
    }
    int candidate = 0;
    unsigned rounds = 0;
    /* the first round starts at once (hb_receive_cb -> start_election); a split vote is retried after a random timeout */
    uint64_t deadline = now_us();
    uint64_t voted_deadline = 0;
    const uint64_t t_start = now_us();
    while (!g_terminate) {
        if (apus_ctl_read(g_rep, &v) != APUS_OK) { LOGT("%s\n", apus_last_error()); return 1; }
        /* (a) somebody won and has adjusted my log: follow.  An announcement below the term I stand in myself is normally
         * ignored -- but a candidacy that timed out without winning gives way to a leader that is newer than the one I lost
         * (`> g_term`): the servers that already follow it no longer answer vote requests (their pumps do not poll them, unlike
         * dare_server.c:1119), so insisting could neither win nor end.  A self-vote helps nobody else win, giving it up is safe;
         * a stale announcement costs one heartbeat timeout and is not accepted twice (g_term moves up to it). */
        const int announced = SID_L(v.leader_sid) && SID_IDX(v.leader_sid) != g_idx && SID_TERM(v.leader_sid) > g_term;
        if (announced && (SID_TERM(v.leader_sid) >= SID_TERM(sid) || (candidate && now_us() >= deadline))) {
            g_term = SID_TERM(v.leader_sid); g_leader_idx = SID_IDX(v.leader_sid);
            apus_ctl_set_sid(g_rep, v.leader_sid);
            if (apus_replica_set_role(g_rep, g_leader_idx, g_term) != APUS_OK || launch_self() != APUS_OK) {
                LOGT("cannot follow p%u: %s\n", (unsigned)g_leader_idx, apus_last_error()); return 1;
            }
            g_removed_mask |= 1u << dead;
            LOGT("[T%llu] follow p%u (log adjusted to end %llu, %llu entries) after %.1f ms\n", (unsigned long long)g_term,
                 (unsigned)g_leader_idx, (unsigned long long)v.adj_end, (unsigned long long)v.adj_count, (now_us() - t_start) / 1e3);
            return 0;
        }
        /* (b) candidate: count the votes (poll_vote_count) */
        if (candidate) {
            unsigned votes = 1;
            uint32_t voters = 0;
            for (unsigned i = 0; i < g_n; i++)
                if (i != g_idx && v.vote_ack[i] != g_log_len) { votes++; voters |= 1u << i; }
            if (votes >= (unsigned)g_n / 2 + 1) {
                /* won.  Give the remaining live servers a moment to answer as well: whoever has not voted by then is
                 * treated as failed (check_failure_count, dare_server.c:1189-1228) */
                uint64_t wait_us = (g_env_hbto_us > 0 ? (uint64_t)g_env_hbto_us : (uint64_t)(10.0 * cfg_hb_period * 1e6)) / 2;
                if (wait_us < cfg_elec_low) wait_us = cfg_elec_low;      /* survivors notice the silence within one hb_timeout of */
                if (wait_us < 5000) wait_us = 5000;                      /* each other; half of it is what their patience leaves me */
                const uint64_t grace = now_us() + wait_us;
                while (now_us() < grace && votes < g_n - 1u) {
                    if (apus_ctl_read(g_rep, &v) != APUS_OK) break;
                    votes = 1; voters = 0;
                    for (unsigned i = 0; i < g_n; i++)
                        if (i != g_idx && v.vote_ack[i] != g_log_len) { votes++; voters |= 1u << i; }
                    usleep(100);
                }
                g_term = SID_TERM(sid); g_leader_idx = g_idx;
                const uint64_t lsid = SID_MAKE(g_term, 1, g_idx);
                apus_ctl_set_sid(g_rep, lsid);
                LOGT("[T%llu] LEADER\n", (unsigned long long)g_term);      /* benchmarks/run.sh:52, reconf_bench.sh grep this */
                if (apus_replica_set_role(g_rep, g_idx, g_term) != APUS_OK) { LOGT("take-over failed: %s\n", apus_last_error()); return 1; }
                for (unsigned i = 0; i < g_n; i++) {
                    if (i == g_idx) continue;
                    if (!(voters & (1u << i))) {
                        apus_replica_disconnect(g_rep, (uint8_t)i);
                        if (g_live_mask & (1u << i)) { g_removed_mask |= 1u << i; LOGT("REMOVE SERVER p%u\n", i); }
                        continue;
                    }
                    uint64_t resent = 0;
                    if (apus_ctl_adjust_follower(g_rep, (uint8_t)i, lsid, &resent) != APUS_OK) {
                        LOGT("log adjustment of p%u failed: %s\n", i, apus_last_error());
                        apus_replica_disconnect(g_rep, (uint8_t)i); g_removed_mask |= 1u << i;
                    } else LOGT("   (p%u: log adjusted, %llu bytes resent)\n", i, (unsigned long long)resent);
                }
                if (launch_self() != APUS_OK) { LOGT("launch: %s\n", apus_last_error()); return 1; }
                LOGT("[T%llu] leading after %.1f ms\n", (unsigned long long)g_term, (now_us() - t_start) / 1e3);
                return 0;
            }
        }
        /* (c) vote requests (poll_vote_requests): the best SID above mine -- with my L flag set, so that I vote once per term */
        uint64_t old_sid = sid | (1ull << 8), best = old_sid;
        int best_i = -1;
        for (unsigned i = 0; i < g_n; i++) {
            if (i == g_idx || v.vote_req[i].sid == 0) continue;
            if (v.vote_req[i].sid <= best) { apus_ctl_clear_vote_request(g_rep, (uint8_t)i); continue; }
            /* up-to-date test: my log must not be ahead of the candidate's (dare_server.c:1640-1652) */
            if (last_term > v.vote_req[i].term || (last_term == v.vote_req[i].term && last_idx > v.vote_req[i].index)) {
                /* my log is better but my term is too low: raise it to improve my own chances (:1663-1676) */
                if (SID_TERM(v.vote_req[i].sid) > SID_TERM(sid)) { sid = SID_MAKE(SID_TERM(v.vote_req[i].sid), 0, g_idx); apus_ctl_set_sid(g_rep, sid); candidate = 0; }
                apus_ctl_clear_vote_request(g_rep, (uint8_t)i);
                continue;
            }
            best = v.vote_req[i].sid; best_i = (int)i;
        }
        if (best_i >= 0) {
            sid = best;                                           /* my vote: SID of the candidate (replicated = written to my words) */
            apus_ctl_set_sid(g_rep, sid);
            candidate = 0;
            LOGT("[T%llu] Vote for p%u\n", (unsigned long long)SID_TERM(sid), (unsigned)SID_IDX(sid));
            apus_ctl_send_vote_ack(g_rep, SID_IDX(sid), commit);
            apus_ctl_clear_vote_request(g_rep, (uint8_t)best_i);
            /* hb_timeout(): the same patience the failure detector has (an environment override included) -- the winner's
             * take-over (grace for late voters, role change, log adjustment) has to fit into it, or this voter stands
             * against a leader that is about to announce itself */
            voted_deadline = now_us() + (g_env_hbto_us > 0 ? (uint64_t)g_env_hbto_us : (uint64_t)(10.0 * cfg_hb_period * 1e6));
            deadline = voted_deadline + random_election_timeout_us();
        }
        /* (d) nobody leads, nobody I voted for made it: stand myself (start_election) */
        if (now_us() >= deadline) {
            if (++rounds > 200) { LOGT("no leader after %u election rounds: giving up\n", rounds); return 1; }
            sid = SID_MAKE(SID_TERM(sid) + 1, 0, g_idx);
            apus_ctl_set_sid(g_rep, sid);
            apus_ctl_reset_votes(g_rep);
            candidate = 1;
            LOGT("[T%llu] Start election\n", (unsigned long long)SID_TERM(sid));
            for (unsigned i = 0; i < g_n; i++)
                if (i != g_idx && i != dead) apus_ctl_send_vote_request(g_rep, (uint8_t)i, sid, last_idx, last_term, cid);
            deadline = now_us() + random_election_timeout_us();
        }
        usleep(50);
    }
    return 1;
}

void *dare_server_init(void *arg)
{
    dare_server_input_t *input = (dare_server_input_t *)arg;
    g_in = *input;
    g_log = input->log ? input->log : stdout;
    free(input);                                       /* dare_server.c:208 */
    signal(SIGINT, int_handler);                       /* dare_server.c:186-187 */

    g_idx = g_in.server_idx; g_n = g_in.group_size;
    /* this thread is created from the interposer's init hook, microseconds before the application's main();
     * let main()'s first instructions (redis: setproctitle rewriting environ) pass before the CUDA runtime
     * starts reading the environment on this thread */
    usleep(10000);
    g_leader_idx = (uint8_t)g_env_leader;
    read_dare_config(g_in.config_path);
    if (g_env_elec_lo > 0) { cfg_elec_low = (uint64_t)g_env_elec_lo; cfg_elec_high = (uint64_t)g_env_elec_hi; }
    if (g_env_hb_us > 0) cfg_hb_period = g_env_hb_us * 1e-6;
    const int joining = (g_in.srv_type == SRV_TYPE_JOIN);
    if (g_in.srv_type != SRV_TYPE_START && !joining) { LOGT("unknown server_type\n"); return NULL; }
    if (g_n < 1 || g_n > APUS_MAX_SERVER_COUNT || g_idx >= g_n) { LOGT("bad group_size/server_idx\n"); return NULL; }

    apus_config_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.struct_size = sizeof cfg;
    int ndev = apus_device_count();
    if (ndev < 1) { LOGT("no CUDA device: the engine has no CPU fallback\n"); return NULL; }
    cfg.device = g_env_gpu >= 0 ? g_env_gpu : (int)(g_idx % (unsigned)ndev);
    if (joining && g_leader_idx == g_idx) g_leader_idx = (uint8_t)((g_idx + 1) % g_n);   /* placeholder until the reply names the leader */
    cfg.server_idx = g_idx; cfg.group_size = g_n; cfg.leader_idx = g_leader_idx;
    cfg.ring_mode = APUS_RING_HOST_MAPPED;
    cfg.flags = APUS_F_EXPLICIT | APUS_F_DEVICE_STATS | APUS_F_AUTOPRUNE | APUS_F_HOST_APPLY;
    cfg.term = g_term;
    g_live_mask = (1u << g_n) - 1u;
    /* heartbeats: the leader kernel beats every hb_period, a follower suspects it after hb_timeout() = 10 periods
     * (dare_server.c:1252-1259) unless the environment says otherwise */
    cfg.hb_period_us = (uint32_t)(cfg_hb_period * 1e6);
    cfg.hb_timeout_us = g_env_hbto_us > 0 ? (uint32_t)g_env_hbto_us : (uint32_t)(10.0 * cfg_hb_period * 1e6);
    if (g_n == 1) { cfg.hb_period_us = 0; cfg.hb_timeout_us = 0; }
    cfg.log_size = g_env_log_size;
    cfg.leader_ctas = 2;
    if (apus_replica_create(&cfg, &g_rep) != APUS_OK) { LOGT("apus_replica_create: %s\n", apus_last_error()); return NULL; }

    if (joining) {
        /* server_type=join: publish my handle, map whoever is there, ask to be let in (join_group) */
        apus_peer_handle_t mine, h;
        char path[512], tmp[560];
        mkdir(g_env_rdv, 0777);
        snprintf(path, sizeof path, "%s/r%u.handle", g_env_rdv, (unsigned)g_idx);
        snprintf(tmp, sizeof tmp, "%s.tmp.%d", path, (int)getpid());
        FILE *hf = fopen(tmp, "wb");
        if (apus_replica_export(g_rep, &mine) != APUS_OK || !hf) { LOGT("cannot publish my peer handle\n"); dare_server_shutdown(); }
        fwrite(&mine, sizeof mine, 1, hf); fclose(hf); rename(tmp, path);
        for (unsigned i = 0; i < g_n; i++)
            if (i != g_idx && read_handle(i, &h) == 0 && apus_replica_connect(g_rep, (uint8_t)i, &h) != APUS_OK)
                LOGT("   (p%u is not reachable: %s)\n", i, apus_last_error());
        if (join_group() != 0 || apus_replica_set_role(g_rep, g_leader_idx, g_term) != APUS_OK || launch_self() != APUS_OK) {
            LOGT("join failed: %s\n", apus_last_error()); dare_server_shutdown();
        }
    } else if (g_env_colocate && g_idx == g_leader_idx && g_n > 1) {
        /* One GPU cannot run the persistent kernels of several PROCESSES at once (contexts are time-sliced), so a
         * single-GPU box can host the followers' replicas inside the leader's process: their kernels ack and follow the
         * commit exactly as anywhere else, only their host side (do_action replay) does not exist.  Measurement aid for
         * the leader-side path through proxy.c; real deployments run one process and one GPU per replica. */
        apus_replica_t *all_r[APUS_MAX_SERVER_COUNT];
        apus_peer_handle_t hs[APUS_MAX_SERVER_COUNT];
        for (unsigned i = 0; i < g_n; i++) {
            if (i == g_idx) { all_r[i] = g_rep; continue; }
            apus_config_t fc = cfg;
            fc.server_idx = (uint8_t)i;
            fc.flags &= ~APUS_F_HOST_APPLY;
            fc.hb_timeout_us = 0;
            if (apus_replica_create(&fc, &all_r[i]) != APUS_OK) { LOGT("colocated follower %u: %s\n", i, apus_last_error()); dare_server_shutdown(); }
        }
        for (unsigned i = 0; i < g_n; i++) apus_replica_export(all_r[i], &hs[i]);
        for (unsigned i = 0; i < g_n; i++)
            for (unsigned j = 0; j < g_n; j++)
                if (i != j && apus_replica_connect(all_r[i], (uint8_t)j, &hs[j]) != APUS_OK) { LOGT("connect: %s\n", apus_last_error()); dare_server_shutdown(); }
        /* followers first in the table: one fused launch */
        apus_replica_t *order[APUS_MAX_SERVER_COUNT];
        unsigned k = 0;
        for (unsigned i = 0; i < g_n; i++) if (i != g_idx) order[k++] = all_r[i];
        order[k++] = g_rep;
        if (apus_replicas_launch(order, (int)k, UINT64_MAX) != APUS_OK) { LOGT("launch: %s\n", apus_last_error()); dare_server_shutdown(); }
        LOGT("followers colocated in the leader's process (single-GPU measurement mode)\n");
    } else {
    apus_peer_handle_t mine, all[APUS_MAX_SERVER_COUNT];
    const char *dir = g_env_rdv;
    if (apus_replica_export(g_rep, &mine) != APUS_OK || rendezvous(dir, &mine, all)) {
        LOGT("peer rendezvous failed in %s\n", dir); dare_server_shutdown();
    }
    for (unsigned i = 0; i < g_n; i++)
        if (i != g_idx && apus_replica_connect(g_rep, (uint8_t)i, &all[i]) != APUS_OK) {
            LOGT("apus_replica_connect(%u): %s\n", i, apus_last_error()); dare_server_shutdown();
        }
    apus_replica_t *rs[1] = { g_rep };
    if (apus_replicas_launch(rs, 1, UINT64_MAX) != APUS_OK) { LOGT("launch: %s\n", apus_last_error()); dare_server_shutdown(); }
    }
    g_tk_type = (uint8_t *)calloc(TK_RING, 1);
    g_started = 1;
    LOGT("replica %u/%u up on GPU %d (leader %u)\n", (unsigned)g_idx, (unsigned)g_n, cfg.device, (unsigned)g_leader_idx);

    g_log_len = cfg.log_size ? cfg.log_size : APUS_LOG_SIZE;
    int elected = 0;
    for (;;) {
        if (g_idx == g_leader_idx) { leader_pump(elected); break; }
        if (!follower_pump(g_log_len)) break;              /* terminated */
        if (elect() != 0) break;                           /* new role, kernel running again */
        elected = 1;
    }
    LOGT("SIGINT detected; shutdown\n");
    dare_server_shutdown();
    return NULL;
}
