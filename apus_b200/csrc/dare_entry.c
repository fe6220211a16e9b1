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
static dare_server_input_t g_in;

#define LOGT(fmt, ...) do { struct timeval _tv; gettimeofday(&_tv, NULL); \
    fprintf(g_log, "[%lu:%06lu] " fmt, (unsigned long)_tv.tv_sec, (unsigned long)_tv.tv_usec, ##__VA_ARGS__); fflush(g_log); } while (0)

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
static unsigned long long g_env_log_size;
static char g_env_rdv[256];
__attribute__((constructor)) static void engine_env_init(void)
{
    const char *s;
    if ((s = getenv("apus_gpu"))) g_env_gpu = atoi(s);
    if ((s = getenv("apus_leader"))) g_env_leader = atoi(s);
    if ((s = getenv("apus_log_size"))) g_env_log_size = strtoull(s, NULL, 0);
    if ((s = getenv("apus_rendezvous"))) snprintf(g_env_rdv, sizeof g_env_rdv, "%s", s);
    else snprintf(g_env_rdv, sizeof g_env_rdv, "/tmp/apus-rdv-%u", (unsigned)getuid());
    if (getenv("apus_segv_trace")) { signal(SIGSEGV, segv_trace); signal(SIGBUS, segv_trace); signal(SIGABRT, segv_trace); }
}

static int csm_like(uint8_t type) { return !(type == APUS_NOOP || type == APUS_CONFIG || type == APUS_HEAD); }

/* ---- leader pump ---------------------------------------------------------------------------- */
#define TK_RING (1u << 20)
static uint8_t *g_tk_type;        /* type of every ticket in flight, indexed by ticket & (TK_RING-1) */

static void leader_pump(void)
{
    uint64_t submitted = 0, applied = 0;
    uint8_t image[64];
    struct apus_tailhead_t pending;                 /* popped from the shared queue, not yet in the engine */
    TAILQ_INIT(&pending);
    if (g_n > 1) {
        /* the election winner's blank CONFIG entry (dare_server.c:1412-1421) */
        uint8_t cid[16];
        memset(cid, 0, sizeof cid);
        cid[8] = g_n;
        uint32_t bm = (1u << g_n) - 1u;
        memcpy(cid + 12, &bm, 4);
        uint64_t t = 0;
        if (apus_submit(g_rep, APUS_CONFIG, 0, 0, cid, 0, &t) != APUS_OK) { LOGT("cannot append CONFIG: %s\n", apus_last_error()); return; }
        g_tk_type[t & (TK_RING - 1)] = APUS_CONFIG;
        submitted = t;
    }
    LOGT("[T%llu] LEADER\n", 1ull);              /* benchmarks/run.sh:52 greps for "] LEADER" */
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
        /* apply_committed_entries, leader branch (dare_server.c:1851-1861, 1951-1952) */
        uint64_t c = apus_committed_tickets(g_rep);
        while (applied < c) {
            applied++;
            if (csm_like(g_tk_type[applied & (TK_RING - 1)]) && g_in.update_state) g_in.update_state(g_in.up_para);
        }
    }
}

/* ---- follower pump: apply_committed_entries, follower branch (dare_server.c:1815-1967) -------- */
static void follower_pump(uint64_t L)
{
    uint64_t apply = 0, next_idx = 0;
    const size_t cap = 1u << 20;
    uint8_t *buf = (uint8_t *)malloc(cap + 65536 + 64);
    uint32_t idle = 0;
    while (!g_terminate) {
        uint64_t off = 0, cnt = 0;
        if (apus_progress(g_rep, &off, &cnt) != APUS_OK) { LOGT("%s\n", apus_last_error()); break; }
        if (off == apply) {
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
    free(buf);
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
    if (g_in.srv_type != SRV_TYPE_START) { LOGT("server_type=join is not supported by the GPU engine yet\n"); return NULL; }
    if (g_n < 1 || g_n > APUS_MAX_SERVER_COUNT || g_idx >= g_n) { LOGT("bad group_size/server_idx\n"); return NULL; }

    apus_config_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.struct_size = sizeof cfg;
    int ndev = apus_device_count();
    if (ndev < 1) { LOGT("no CUDA device: the engine has no CPU fallback\n"); return NULL; }
    cfg.device = g_env_gpu >= 0 ? g_env_gpu : (int)(g_idx % (unsigned)ndev);
    cfg.server_idx = g_idx; cfg.group_size = g_n; cfg.leader_idx = g_leader_idx;
    cfg.ring_mode = APUS_RING_HOST_MAPPED;
    cfg.flags = APUS_F_EXPLICIT | APUS_F_DEVICE_STATS | APUS_F_AUTOPRUNE | APUS_F_HOST_APPLY;
    cfg.term = 1;                                      /* term of a clean first election */
    cfg.log_size = g_env_log_size;
    cfg.leader_ctas = 2;
    if (apus_replica_create(&cfg, &g_rep) != APUS_OK) { LOGT("apus_replica_create: %s\n", apus_last_error()); return NULL; }

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
    g_tk_type = (uint8_t *)calloc(TK_RING, 1);
    g_started = 1;
    LOGT("replica %u/%u up on GPU %d (leader %u)\n", (unsigned)g_idx, (unsigned)g_n, cfg.device, (unsigned)g_leader_idx);

    if (g_idx == g_leader_idx) leader_pump();
    else follower_pump(cfg.log_size ? cfg.log_size : APUS_LOG_SIZE);
    LOGT("SIGINT detected; shutdown\n");
    dare_server_shutdown();
    return NULL;
}
