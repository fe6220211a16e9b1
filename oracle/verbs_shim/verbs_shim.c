/*
 * oracle/verbs_shim/verbs_shim.c -- TEST INFRASTRUCTURE ONLY (see infiniband/verbs.h in this directory).
 *
 * "reference-on-shim" (CPU baseline option 2): the reference's own RDMA code runs unmodified; this
 * file is the NIC.  One replica = one process.
 *
 *   RC queue pairs   RDMA WRITE / READ  -> process_vm_writev / process_vm_readv into the peer process, executed
 *                    synchronously inside ibv_post_send; the completion (if signalled, or on error) is queued
 *                    on the send CQ at once.  Before every operation the REMOTE queue pair is examined (its
 *                    published record is read from the peer): it must be an RC QP in RTR/RTS, connected back
 *                    to this QP, with rq_psn equal to this side's sq_psn -- otherwise the operation fails with
 *                    IBV_WC_RETRY_EXC_ERR and the local QP enters the error state, which is what a real HCA
 *                    reports when the responder has revoked access (DARE's log-access fencing relies on it,
 *                    dare_ibv_rc.c:2150-2250).  The remote key / address range is checked against the peer's
 *                    published memory regions (IBV_WC_REM_ACCESS_ERR).
 *   UD queue pairs   SEND -> one Unix datagram to the socket of the destination port ($APUS_SHIM_DIR/port.<lid>);
 *                    a destination LID >= 0xC000 is a multicast group: the datagram goes to every port in the
 *                    directory, the sender included (the reference filters its own LID, dare_ibv_ud.c:811).
 *                    Receives are pulled from the socket when the receive CQ is polled; 40 bytes of GRH precede
 *                    the payload in the posted buffer, as on a real UD QP.
 *   addressing       port LID = $APUS_SHIM_ID (default: $server_idx + 1); GID = fe80::a9:<lid>; the GID is what
 *                    routes RC traffic (RoCE style), so the reference's hostname-derived "unique slid" is harmless.
 *
 * Latency of the emulated wire: one or two system calls (about a microsecond); there is no NIC, no retransmission
 * and no loss.  That makes the timing an UPPER bound on what the reference's software path can do.
 *
 *   APUS_SHIM_TRANSPORT=shm   the zero-latency variant BASELINE.md s2 planned: the published QP / MR record lives in a
 *                    memfd every peer maps, and a memory region of 1 MiB or more (the log) is re-backed IN PLACE by a
 *                    memfd (same virtual address, contents kept) that the peers map too -- an RDMA WRITE / READ into it
 *                    is a memcpy + store fence, the responder-QP check is a load; no system call on the replicate path.
 *                    Smaller regions (control data: heartbeats, votes) keep the process_vm path, which is also what
 *                    notices a dead peer (ESRCH); for mapped regions the peer's pid is probed every 64 operations.
 */
#define _GNU_SOURCE
#include <dirent.h>
#include <errno.h>
#include <fcntl.h>
#include <signal.h>
#include <sys/mman.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/prctl.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/uio.h>
#include <sys/un.h>
#include <unistd.h>

#include "infiniband/verbs.h"

#define SHIM_MAX_QP 256
#define SHIM_MAX_MR 4096
#define SHIM_MAX_PEER 64
#define SHIM_MAX_INLINE 512
#define SHIM_MCAST_LID 0xC000

/* what a peer may read about this process (process_vm_readv) */
typedef struct { uint32_t qpn, type, state, dest_qpn, rq_psn, sq_psn; } pub_qp_t;
typedef struct { uint32_t rkey, access; uint64_t addr, len; int32_t shm_fd; uint32_t shm_gen; uint64_t shm_base, shm_len; } pub_mr_t;
typedef struct { pub_qp_t qp[SHIM_MAX_QP]; pub_mr_t mr[SHIM_MAX_MR]; } pub_t;
static pub_t g_pub_static __attribute__((aligned(4096)));
static pub_t *g_pubp = &g_pub_static;        /* shm transport: a memfd mapping the peers share */
#define g_pub (*g_pubp)
static size_t g_shm_min = 1u << 20;          /* regions this large are re-backed by a memfd the peers map (APUS_SHIM_SHM_MIN) */
#define SHIM_MAX_SEG 8

typedef struct { struct ibv_wc *ring; int cap, head, tail; struct ibv_qp *ud_recv_qp; } shim_cq_t;
typedef struct { uint64_t wr_id; uint64_t addr; uint32_t len; } shim_rwr_t;
typedef struct {
    int idx;
    struct ibv_qp_attr attr;
    struct ibv_qp_init_attr init;
    int mcast;
    shim_rwr_t *rq; int rq_cap, rq_head, rq_tail;
} shim_qp_t;

typedef struct { uint32_t rkey, gen; int32_t fd; uint8_t *map; uint64_t base, len; } peer_seg_t;
typedef struct {
    int known; pid_t pid; uint64_t pub_addr; pub_mr_t mr_cache[64];
    int pub_fd; pub_t *pub_map;              /* shm transport: the peer's public record, mapped */
    peer_seg_t seg[SHIM_MAX_SEG];            /* ... and its large regions */
    uint32_t ops;
} peer_t;

static struct ibv_device g_dev = { "apus_shim0" };
static struct ibv_device *g_dev_list[2] = { &g_dev, NULL };
static struct ibv_context g_ctx = { &g_dev, -1 };
static uint16_t g_lid;
static char g_dir[200];
static int g_sock = -1;
static peer_t g_peer[SHIM_MAX_PEER];
static struct ibv_qp *g_qps[SHIM_MAX_QP];
static uint32_t g_next_handle = 1;
static int g_trace;
static int g_shm;                            /* APUS_SHIM_TRANSPORT=shm */
static uint64_t g_ops_memcpy, g_ops_vm;      /* RC operations carried by a memcpy / by process_vm_* (shim_transport_stats) */
static int g_pub_fd = -1;
static int g_mr_fd[SHIM_MAX_MR];

#define TRACE(...) do { if (g_trace) { fprintf(stderr, "[shim %u] ", (unsigned)g_lid); fprintf(stderr, __VA_ARGS__); } } while (0)

static void gid_of(uint16_t lid, union ibv_gid *gid)
{
    memset(gid, 0, sizeof *gid);
    gid->raw[0] = 0xfe; gid->raw[1] = 0x80; gid->raw[13] = 0xa9; gid->raw[14] = (uint8_t)(lid >> 8); gid->raw[15] = (uint8_t)lid;
}
static uint16_t lid_of_gid(const union ibv_gid *gid) { return (uint16_t)((gid->raw[14] << 8) | gid->raw[15]); }

static void shim_init(void)
{
    if (g_lid) return;
    const char *s;
    g_trace = getenv("APUS_SHIM_TRACE") != NULL;
    if ((s = getenv("APUS_SHIM_ID"))) g_lid = (uint16_t)atoi(s);
    else if ((s = getenv("server_idx"))) g_lid = (uint16_t)(atoi(s) + 1);
    else g_lid = (uint16_t)(1 + (getpid() % (SHIM_MAX_PEER - 1)));
    if (g_lid == 0 || g_lid >= SHIM_MAX_PEER) { fprintf(stderr, "verbs shim: bad APUS_SHIM_ID\n"); exit(1); }
    snprintf(g_dir, sizeof g_dir, "%s", (s = getenv("APUS_SHIM_DIR")) ? s : "/tmp/apus-verbs-shim");
    mkdir(g_dir, 0777);
    if ((s = getenv("APUS_SHIM_TRANSPORT")) && !strcmp(s, "shm")) {
        int fd = memfd_create("apus-shim-pub", MFD_CLOEXEC);
        void *m = MAP_FAILED;
        if (fd >= 0 && ftruncate(fd, sizeof(pub_t)) == 0) m = mmap(NULL, sizeof(pub_t), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        if (m == MAP_FAILED) { perror("verbs shim: shm transport unavailable, using process_vm"); if (fd >= 0) close(fd); }
        else { g_pubp = m; g_pub_fd = fd; g_shm = 1; }
        if ((s = getenv("APUS_SHIM_SHM_MIN"))) g_shm_min = (size_t)strtoull(s, NULL, 0);
    }
    prctl(PR_SET_PTRACER, PR_SET_PTRACER_ANY, 0, 0, 0);           /* let the peers write into this process */
    /* UD port */
    struct sockaddr_un sa;
    memset(&sa, 0, sizeof sa);
    sa.sun_family = AF_UNIX;
    snprintf(sa.sun_path, sizeof sa.sun_path, "%s/port.%u", g_dir, (unsigned)g_lid);
    unlink(sa.sun_path);
    g_sock = socket(AF_UNIX, SOCK_DGRAM | SOCK_NONBLOCK | SOCK_CLOEXEC, 0);
    int sz = 4 << 20;
    setsockopt(g_sock, SOL_SOCKET, SO_RCVBUF, &sz, sizeof sz);
    setsockopt(g_sock, SOL_SOCKET, SO_SNDBUF, &sz, sizeof sz);
    if (g_sock < 0 || bind(g_sock, (struct sockaddr *)&sa, sizeof sa)) { perror("verbs shim: bind"); exit(1); }
    /* publish pid + address of the public record */
    char path[256], tmp[300];
    snprintf(path, sizeof path, "%s/proc.%u", g_dir, (unsigned)g_lid);
    snprintf(tmp, sizeof tmp, "%s.tmp", path);
    FILE *f = fopen(tmp, "w");
    if (!f) { perror("verbs shim: proc file"); exit(1); }
    fprintf(f, "%d %llu %d\n", (int)getpid(), (unsigned long long)(uintptr_t)&g_pub, g_pub_fd);
    fclose(f);
    rename(tmp, path);
}

static void *map_peer_fd(pid_t pid, int fd, size_t len)
{
    char path[64];
    snprintf(path, sizeof path, "/proc/%d/fd/%d", (int)pid, fd);
    int h = open(path, O_RDWR | O_CLOEXEC);
    if (h < 0) return NULL;
    void *m = mmap(NULL, len, PROT_READ | PROT_WRITE, MAP_SHARED, h, 0);
    close(h);
    return m == MAP_FAILED ? NULL : m;
}
static void peer_unmap(peer_t *p)
{
    if (p->pub_map) munmap(p->pub_map, sizeof(pub_t));
    for (int i = 0; i < SHIM_MAX_SEG; i++) if (p->seg[i].map) munmap(p->seg[i].map, p->seg[i].len);
    p->pub_map = NULL;
    memset(p->seg, 0, sizeof p->seg);
}
/* the peer's mapping of a large region of `p`, or NULL (then the caller uses process_vm) */
static uint8_t *peer_seg(peer_t *p, const pub_mr_t *m)
{
    if (!p->pub_map || m->shm_fd < 0 || !m->shm_len) return NULL;
    peer_seg_t *freep = NULL;
    for (int i = 0; i < SHIM_MAX_SEG; i++) {
        peer_seg_t *g = &p->seg[i];
        if (g->map && g->rkey == m->rkey) {
            if (g->gen == m->shm_gen && g->fd == m->shm_fd && g->base == m->shm_base && g->len == m->shm_len) return g->map;
            munmap(g->map, g->len); memset(g, 0, sizeof *g);  /* the region was registered anew */
        }
        if (!g->map && !freep) freep = g;
    }
    if (!freep) return NULL;
    void *mm = map_peer_fd(p->pid, m->shm_fd, m->shm_len);
    if (!mm) return NULL;
    freep->rkey = m->rkey; freep->gen = m->shm_gen; freep->fd = m->shm_fd; freep->map = mm; freep->base = m->shm_base; freep->len = m->shm_len;
    return mm;
}

static peer_t *peer_of(uint16_t lid, int refresh)
{
    if (lid == 0 || lid >= SHIM_MAX_PEER) return NULL;
    peer_t *p = &g_peer[lid];
    if (p->known && !refresh) return p;
    char path[256];
    snprintf(path, sizeof path, "%s/proc.%u", g_dir, (unsigned)lid);
    FILE *f = fopen(path, "r");
    if (!f) return NULL;
    int pid, pfd = -1; unsigned long long a;
    int got = fscanf(f, "%d %llu %d", &pid, &a, &pfd);
    fclose(f);
    if (got < 2) return NULL;
    peer_unmap(p);
    memset(p, 0, sizeof *p);
    p->known = 1; p->pid = pid; p->pub_addr = a; p->pub_fd = got == 3 ? pfd : -1;
    if (g_shm && p->pub_fd >= 0) {
        void *m = map_peer_fd(pid, p->pub_fd, sizeof(pub_t));
        if (m) p->pub_map = m;                                 /* else: this peer is reached through process_vm */
    }
    return p;
}

static int peer_read(peer_t *p, uint64_t raddr, void *dst, size_t n)
{
    struct iovec l = { dst, n }, r = { (void *)(uintptr_t)raddr, n };
    return process_vm_readv(p->pid, &l, 1, &r, 1, 0) == (ssize_t)n ? 0 : -1;
}
static int peer_write(peer_t *p, uint64_t raddr, const void *src, size_t n)
{
    struct iovec l = { (void *)src, n }, r = { (void *)(uintptr_t)raddr, n };
    return process_vm_writev(p->pid, &l, 1, &r, 1, 0) == (ssize_t)n ? 0 : -1;
}

/* ---- devices ----------------------------------------------------------------------------------------------- */
struct ibv_device **ibv_get_device_list(int *n) { shim_init(); if (n) *n = 1; return g_dev_list; }
void ibv_free_device_list(struct ibv_device **l) { (void)l; }
const char *ibv_get_device_name(struct ibv_device *d) { return d ? d->name : "?"; }
struct ibv_context *ibv_open_device(struct ibv_device *d) { (void)d; shim_init(); return &g_ctx; }
int ibv_close_device(struct ibv_context *c)
{
    (void)c;
    char path[256];
    snprintf(path, sizeof path, "%s/port.%u", g_dir, (unsigned)g_lid); unlink(path);
    snprintf(path, sizeof path, "%s/proc.%u", g_dir, (unsigned)g_lid); unlink(path);
    return 0;
}
int ibv_query_device(struct ibv_context *c, struct ibv_device_attr *a)
{
    (void)c;
    memset(a, 0, sizeof *a);
    snprintf(a->fw_ver, sizeof a->fw_ver, "shim");
    a->max_qp = SHIM_MAX_QP; a->max_qp_wr = 1024; a->max_sge = 4; a->max_cq = 64; a->max_cqe = 65536; a->max_mr = SHIM_MAX_MR;
    a->max_pd = 16; a->max_qp_rd_atom = 16; a->max_res_rd_atom = 16; a->max_qp_init_rd_atom = 16;
    a->atomic_cap = IBV_ATOMIC_NONE; a->max_mcast_grp = 16; a->max_mcast_qp_attach = 16; a->max_ah = 1024;
    a->max_srq = 0; a->max_pkeys = 1; a->phys_port_cnt = 1; a->max_mr_size = ~0ull;
    return 0;
}
int ibv_query_port(struct ibv_context *c, uint8_t port, struct ibv_port_attr *a)
{
    (void)c; (void)port;
    memset(a, 0, sizeof *a);
    a->state = IBV_PORT_ACTIVE; a->max_mtu = IBV_MTU_4096; a->active_mtu = IBV_MTU_4096; a->gid_tbl_len = 1;
    a->pkey_tbl_len = 1; a->lid = g_lid; a->link_layer = IBV_LINK_LAYER_ETHERNET; a->max_msg_sz = 1u << 30;
    return 0;
}
int ibv_query_gid(struct ibv_context *c, uint8_t port, int index, union ibv_gid *gid) { (void)c; (void)port; (void)index; gid_of(g_lid, gid); return 0; }
int ibv_query_pkey(struct ibv_context *c, uint8_t port, int index, uint16_t *pkey) { (void)c; (void)port; (void)index; *pkey = 0xFFFF; return 0; }

struct ibv_pd *ibv_alloc_pd(struct ibv_context *c) { struct ibv_pd *pd = calloc(1, sizeof *pd); pd->context = c; pd->handle = g_next_handle++; return pd; }
int ibv_dealloc_pd(struct ibv_pd *pd) { free(pd); return 0; }

/* ---- memory regions ------------------------------------------------------------------------------------------ */
/* shm transport: put the pages of [addr, addr+length) on a memfd, at the same virtual address and with the same
 * contents, so that peers can map them.  The page-rounded range is what moves; for the one region this applies to (the
 * log: a malloc of 64 MiB, i.e. its own anonymous mapping) the rounding adds the chunk header and the tail slack only.
 * On any failure the region simply stays process_vm-only. */
static void shm_back_region(int i, void *addr, size_t length)
{
    static uint32_t gen;
    uintptr_t base = (uintptr_t)addr & ~(uintptr_t)4095, end = ((uintptr_t)addr + length + 4095) & ~(uintptr_t)4095;
    size_t sz = end - base;
    int fd = memfd_create("apus-shim-mr", MFD_CLOEXEC);
    if (fd < 0) return;
    void *tmp = MAP_FAILED;
    if (ftruncate(fd, (off_t)sz) == 0) tmp = mmap(NULL, sz, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    if (tmp == MAP_FAILED) { close(fd); return; }
    memcpy(tmp, (void *)base, sz);
    munmap(tmp, sz);
    if (mmap((void *)base, sz, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_FIXED, fd, 0) == MAP_FAILED) { close(fd); return; }
    g_mr_fd[i] = fd;
    g_pub.mr[i].shm_base = base; g_pub.mr[i].shm_len = sz; g_pub.mr[i].shm_gen = ++gen; g_pub.mr[i].shm_fd = fd;
    TRACE("region %d: %zu bytes at %p re-backed by memfd %d\n", i, sz, (void *)base, fd);
}

struct ibv_mr *ibv_reg_mr(struct ibv_pd *pd, void *addr, size_t length, int access)
{
    for (int i = 0; i < SHIM_MAX_MR; i++) {
        if (g_pub.mr[i].rkey) continue;
        struct ibv_mr *mr = calloc(1, sizeof *mr);
        mr->context = pd->context; mr->pd = pd; mr->addr = addr; mr->length = length; mr->handle = (uint32_t)i;
        mr->lkey = mr->rkey = ((uint32_t)g_lid << 16) | (uint32_t)(i + 1);
        g_pub.mr[i].addr = (uint64_t)(uintptr_t)addr; g_pub.mr[i].len = length; g_pub.mr[i].access = (uint32_t)access;
        g_pub.mr[i].shm_fd = -1; g_pub.mr[i].shm_base = g_pub.mr[i].shm_len = 0; g_mr_fd[i] = -1;
        if (g_shm && length >= g_shm_min) shm_back_region(i, addr, length);
        __atomic_store_n(&g_pub.mr[i].rkey, mr->rkey, __ATOMIC_RELEASE);
        return mr;
    }
    errno = ENOMEM;
    return NULL;
}
int ibv_dereg_mr(struct ibv_mr *mr)
{
    if (!mr) return EINVAL;
    __atomic_store_n(&g_pub.mr[mr->handle].rkey, 0, __ATOMIC_RELEASE);
    /* the pages stay memfd-backed until the owner frees them (free() of a large chunk unmaps); peers keep their own mapping */
    if (g_mr_fd[mr->handle] >= 0) { g_pub.mr[mr->handle].shm_fd = -1; close(g_mr_fd[mr->handle]); g_mr_fd[mr->handle] = -1; }
    free(mr);
    return 0;
}

/* ---- completion queues ----------------------------------------------------------------------------------------- */
struct ibv_cq *ibv_create_cq(struct ibv_context *c, int cqe, void *cq_context, void *channel, int comp_vector)
{
    (void)channel; (void)comp_vector;
    struct ibv_cq *cq = calloc(1, sizeof *cq);
    shim_cq_t *s = calloc(1, sizeof *s);
    s->cap = cqe + 64;
    s->ring = calloc((size_t)s->cap, sizeof(struct ibv_wc));
    cq->context = c; cq->cq_context = cq_context; cq->cqe = cqe; cq->handle = g_next_handle++; cq->shim = s;
    return cq;
}
int ibv_destroy_cq(struct ibv_cq *cq) { if (cq) { shim_cq_t *s = cq->shim; free(s->ring); free(s); free(cq); } return 0; }

static void cq_push(struct ibv_cq *cq, const struct ibv_wc *wc)
{
    shim_cq_t *s = cq->shim;
    int next = (s->tail + 1) % s->cap;
    if (next == s->head) {                        /* overrun: grow (a real CQ would raise a fatal async event) */
        int ncap = s->cap * 2, n = 0;
        struct ibv_wc *nr = calloc((size_t)ncap, sizeof *nr);
        while (s->head != s->tail) { nr[n++] = s->ring[s->head]; s->head = (s->head + 1) % s->cap; }
        free(s->ring); s->ring = nr; s->cap = ncap; s->head = 0; s->tail = n;
        next = s->tail + 1;
    }
    s->ring[s->tail] = *wc;
    s->tail = next;
}

typedef struct { uint16_t slid; uint16_t pad; uint32_t src_qpn, dst_qpn; uint8_t sgid[16], dgid[16]; } dgram_hdr_t;

/* pull datagrams from the port into the posted receive buffers of the UD QP that feeds this CQ */
static void ud_pump(struct ibv_cq *cq)
{
    shim_cq_t *s = cq->shim;
    struct ibv_qp *qp = s->ud_recv_qp;
    if (!qp) return;
    shim_qp_t *q = qp->shim;
    if (q->attr.qp_state < IBV_QPS_RTR || q->attr.qp_state == IBV_QPS_ERR) return;
    static uint8_t buf[sizeof(dgram_hdr_t) + 8192];
    /* Polling an empty completion queue is a memory read on a real HCA; here it would be a recv() system call, and the
     * reference polls its UD queue in every turn of its event loop (dare_server.c:1020 poll_ud) -- about half of a
     * replica's CPU time went into EAGAIN.  After an empty poll the socket is left alone for ~15 us (UD carries the
     * start-up handshake and join requests only; nothing on the replication path waits for it). */
    static uint64_t quiet_since;
    const uint64_t now = __builtin_ia32_rdtsc();
    if (quiet_since && now - quiet_since < 40000u) return;
    quiet_since = 0;
    for (int budget = 0; budget < 16; budget++) {
        if (q->rq_head == q->rq_tail) return;                     /* no receive posted: leave it in the socket */
        ssize_t n = recv(g_sock, buf, sizeof buf, MSG_DONTWAIT);
        if (n < (ssize_t)sizeof(dgram_hdr_t)) { quiet_since = now | 1; return; }
        dgram_hdr_t *h = (dgram_hdr_t *)buf;
        if (h->dst_qpn != 0xFFFFFF && h->dst_qpn != qp->qp_num) continue;
        if (h->dst_qpn == 0xFFFFFF && !q->mcast) continue;
        shim_rwr_t w = q->rq[q->rq_head];
        q->rq_head = (q->rq_head + 1) % q->rq_cap;
        size_t pay = (size_t)n - sizeof *h;
        struct ibv_wc wc;
        memset(&wc, 0, sizeof wc);
        wc.wr_id = w.wr_id; wc.opcode = IBV_WC_RECV; wc.qp_num = qp->qp_num; wc.src_qp = h->src_qpn; wc.slid = h->slid;
        wc.wc_flags = IBV_WC_GRH;
        if (pay + 40 > w.len) { wc.status = IBV_WC_LOC_LEN_ERR; cq_push(cq, &wc); continue; }
        struct ibv_grh grh;
        memset(&grh, 0, sizeof grh);
        grh.version_tclass_flow = 0x60; grh.paylen = (uint16_t)((pay >> 8) | (pay << 8)); grh.next_hdr = 0x1b; grh.hop_limit = 0xff;
        memcpy(grh.sgid.raw, h->sgid, 16); memcpy(grh.dgid.raw, h->dgid, 16);
        memcpy((void *)(uintptr_t)w.addr, &grh, 40);
        memcpy((uint8_t *)(uintptr_t)w.addr + 40, buf + sizeof *h, pay);
        wc.status = IBV_WC_SUCCESS; wc.byte_len = (uint32_t)(pay + 40);
        cq_push(cq, &wc);
    }
}

int ibv_poll_cq(struct ibv_cq *cq, int num_entries, struct ibv_wc *wc)
{
    shim_cq_t *s = cq->shim;
    if (s->ud_recv_qp && s->head == s->tail) ud_pump(cq);
    int n = 0;
    while (n < num_entries && s->head != s->tail) { wc[n++] = s->ring[s->head]; s->head = (s->head + 1) % s->cap; }
    return n;
}

/* ---- queue pairs --------------------------------------------------------------------------------------------------- */
struct ibv_qp *ibv_create_qp(struct ibv_pd *pd, struct ibv_qp_init_attr *ia)
{
    if (ia->cap.max_inline_data > SHIM_MAX_INLINE) { errno = EINVAL; return NULL; }
    if (ia->qp_type != IBV_QPT_RC && ia->qp_type != IBV_QPT_UD) { errno = ENOSYS; return NULL; }
    int idx = -1;
    for (int i = 0; i < SHIM_MAX_QP; i++) if (!g_qps[i]) { idx = i; break; }
    if (idx < 0) { errno = ENOMEM; return NULL; }
    struct ibv_qp *qp = calloc(1, sizeof *qp);
    shim_qp_t *q = calloc(1, sizeof *q);
    q->idx = idx; q->init = *ia;
    q->rq_cap = (int)ia->cap.max_recv_wr + 2;
    q->rq = calloc((size_t)q->rq_cap, sizeof *q->rq);
    qp->context = pd->context; qp->pd = pd; qp->send_cq = ia->send_cq; qp->recv_cq = ia->recv_cq; qp->qp_context = ia->qp_context;
    qp->handle = g_next_handle++; qp->qp_num = ((uint32_t)g_lid << 12) | (uint32_t)idx | 0x100000u; qp->qp_type = ia->qp_type;
    qp->state = IBV_QPS_RESET; qp->shim = q;
    g_qps[idx] = qp;
    pub_qp_t *p = &g_pub.qp[idx];
    memset(p, 0, sizeof *p);
    p->type = (uint32_t)ia->qp_type; p->state = IBV_QPS_RESET;
    __atomic_store_n(&p->qpn, qp->qp_num, __ATOMIC_RELEASE);
    if (ia->qp_type == IBV_QPT_UD && ia->recv_cq) ((shim_cq_t *)ia->recv_cq->shim)->ud_recv_qp = qp;
    return qp;
}
int ibv_destroy_qp(struct ibv_qp *qp)
{
    if (!qp) return EINVAL;
    shim_qp_t *q = qp->shim;
    if (qp->qp_type == IBV_QPT_UD && qp->recv_cq && ((shim_cq_t *)qp->recv_cq->shim)->ud_recv_qp == qp)
        ((shim_cq_t *)qp->recv_cq->shim)->ud_recv_qp = NULL;
    memset(&g_pub.qp[q->idx], 0, sizeof(pub_qp_t));
    g_qps[q->idx] = NULL;
    free(q->rq); free(q); free(qp);
    return 0;
}
int ibv_modify_qp(struct ibv_qp *qp, struct ibv_qp_attr *a, int mask)
{
    shim_qp_t *q = qp->shim;
    pub_qp_t *p = &g_pub.qp[q->idx];
    if (mask & IBV_QP_ACCESS_FLAGS) q->attr.qp_access_flags = a->qp_access_flags;
    if (mask & IBV_QP_PKEY_INDEX) q->attr.pkey_index = a->pkey_index;
    if (mask & IBV_QP_PORT) q->attr.port_num = a->port_num;
    if (mask & IBV_QP_QKEY) q->attr.qkey = a->qkey;
    if (mask & IBV_QP_AV) q->attr.ah_attr = a->ah_attr;
    if (mask & IBV_QP_PATH_MTU) q->attr.path_mtu = a->path_mtu;
    if (mask & IBV_QP_TIMEOUT) q->attr.timeout = a->timeout;
    if (mask & IBV_QP_RETRY_CNT) q->attr.retry_cnt = a->retry_cnt;
    if (mask & IBV_QP_RNR_RETRY) q->attr.rnr_retry = a->rnr_retry;
    if (mask & IBV_QP_RQ_PSN) { q->attr.rq_psn = a->rq_psn & 0xFFFFFF; p->rq_psn = q->attr.rq_psn; }
    if (mask & IBV_QP_SQ_PSN) { q->attr.sq_psn = a->sq_psn & 0xFFFFFF; p->sq_psn = q->attr.sq_psn; }
    if (mask & IBV_QP_MAX_QP_RD_ATOMIC) q->attr.max_rd_atomic = a->max_rd_atomic;
    if (mask & IBV_QP_MAX_DEST_RD_ATOMIC) q->attr.max_dest_rd_atomic = a->max_dest_rd_atomic;
    if (mask & IBV_QP_MIN_RNR_TIMER) q->attr.min_rnr_timer = a->min_rnr_timer;
    if (mask & IBV_QP_DEST_QPN) { q->attr.dest_qp_num = a->dest_qp_num; p->dest_qpn = a->dest_qp_num; }
    if (mask & IBV_QP_STATE) {
        enum ibv_qp_state from = q->attr.qp_state, to = a->qp_state;
        int ok = to == IBV_QPS_RESET || to == IBV_QPS_ERR || (from == IBV_QPS_RESET && to == IBV_QPS_INIT) ||
                 (from == IBV_QPS_INIT && (to == IBV_QPS_RTR || to == IBV_QPS_INIT)) ||
                 (from == IBV_QPS_RTR && to == IBV_QPS_RTS) || (from == IBV_QPS_RTS && to == IBV_QPS_RTS);
        if (!ok) return EINVAL;
        if (to == IBV_QPS_RESET) {
            memset(&q->attr, 0, sizeof q->attr);
            q->rq_head = q->rq_tail = 0;
            p->dest_qpn = p->rq_psn = p->sq_psn = 0;
        }
        q->attr.qp_state = q->attr.cur_qp_state = to;
        qp->state = to;
        __atomic_store_n(&p->state, (uint32_t)to, __ATOMIC_RELEASE);
        TRACE("qp %x -> state %d (dest %x rq_psn %u sq_psn %u)\n", qp->qp_num, (int)to, p->dest_qpn, p->rq_psn, p->sq_psn);
    }
    return 0;
}
int ibv_query_qp(struct ibv_qp *qp, struct ibv_qp_attr *a, int mask, struct ibv_qp_init_attr *ia)
{
    (void)mask;
    shim_qp_t *q = qp->shim;
    *a = q->attr;
    if (ia) *ia = q->init;
    return 0;
}

struct ibv_ah *ibv_create_ah(struct ibv_pd *pd, struct ibv_ah_attr *attr)
{
    struct ibv_ah *ah = calloc(1, sizeof *ah);
    ah->context = pd->context; ah->pd = pd; ah->handle = g_next_handle++; ah->attr = *attr;
    return ah;
}
int ibv_destroy_ah(struct ibv_ah *ah) { free(ah); return 0; }
int ibv_attach_mcast(struct ibv_qp *qp, const union ibv_gid *gid, uint16_t lid) { (void)gid; (void)lid; ((shim_qp_t *)qp->shim)->mcast = 1; return 0; }
int ibv_detach_mcast(struct ibv_qp *qp, const union ibv_gid *gid, uint16_t lid) { (void)gid; (void)lid; ((shim_qp_t *)qp->shim)->mcast = 0; return 0; }

int ibv_post_recv(struct ibv_qp *qp, struct ibv_recv_wr *wr, struct ibv_recv_wr **bad)
{
    shim_qp_t *q = qp->shim;
    for (; wr; wr = wr->next) {
        int next = (q->rq_tail + 1) % q->rq_cap;
        if (next == q->rq_head || wr->num_sge != 1) { if (bad) *bad = wr; return ENOMEM; }
        q->rq[q->rq_tail].wr_id = wr->wr_id; q->rq[q->rq_tail].addr = wr->sg_list[0].addr; q->rq[q->rq_tail].len = wr->sg_list[0].length;
        q->rq_tail = next;
    }
    return 0;
}

static void complete(struct ibv_qp *qp, struct ibv_send_wr *wr, enum ibv_wc_status st, uint32_t len)
{
    struct ibv_wc wc;
    memset(&wc, 0, sizeof wc);
    wc.wr_id = wr->wr_id; wc.status = st; wc.qp_num = qp->qp_num; wc.byte_len = len;
    wc.opcode = wr->opcode == IBV_WR_RDMA_READ ? IBV_WC_RDMA_READ : wr->opcode == IBV_WR_SEND ? IBV_WC_SEND : IBV_WC_RDMA_WRITE;
    cq_push(qp->send_cq, &wc);
}

static void qp_to_error(struct ibv_qp *qp)
{
    shim_qp_t *q = qp->shim;
    q->attr.qp_state = q->attr.cur_qp_state = IBV_QPS_ERR; qp->state = IBV_QPS_ERR;
    __atomic_store_n(&g_pub.qp[q->idx].state, (uint32_t)IBV_QPS_ERR, __ATOMIC_RELEASE);
}

static int ud_send_one(uint16_t dlid, const void *msg, size_t len)
{
    struct sockaddr_un sa;
    memset(&sa, 0, sizeof sa);
    sa.sun_family = AF_UNIX;
    snprintf(sa.sun_path, sizeof sa.sun_path, "%s/port.%u", g_dir, (unsigned)dlid);
    for (int tries = 0; tries < 200; tries++) {
        if (sendto(g_sock, msg, len, 0, (struct sockaddr *)&sa, sizeof sa) >= 0) return 0;
        if (errno != EAGAIN && errno != ENOBUFS) return -1;      /* no such port: UD is unreliable, the message is lost */
        usleep(50);
    }
    return -1;
}

static int post_ud(struct ibv_qp *qp, struct ibv_send_wr *wr)
{
    static uint8_t buf[sizeof(dgram_hdr_t) + 8192];
    dgram_hdr_t *h = (dgram_hdr_t *)buf;
    size_t len = 0;
    for (int i = 0; i < wr->num_sge; i++) {
        if (len + wr->sg_list[i].length > 8192) return EINVAL;
        memcpy(buf + sizeof *h + len, (void *)(uintptr_t)wr->sg_list[i].addr, wr->sg_list[i].length);
        len += wr->sg_list[i].length;
    }
    struct ibv_ah *ah = wr->wr.ud.ah;
    if (!ah) return EINVAL;
    union ibv_gid me;
    gid_of(g_lid, &me);
    memset(h, 0, sizeof *h);
    h->slid = g_lid; h->src_qpn = qp->qp_num;
    memcpy(h->sgid, me.raw, 16); memcpy(h->dgid, ah->attr.grh.dgid.raw, 16);
    if (ah->attr.dlid >= SHIM_MCAST_LID) {
        h->dst_qpn = 0xFFFFFF;
        DIR *d = opendir(g_dir);
        struct dirent *e;
        while (d && (e = readdir(d))) {
            unsigned lid;
            if (sscanf(e->d_name, "port.%u", &lid) == 1) ud_send_one((uint16_t)lid, buf, sizeof *h + len);
        }
        if (d) closedir(d);
    } else {
        h->dst_qpn = wr->wr.ud.remote_qpn;
        uint16_t dlid = ah->attr.is_global ? lid_of_gid(&ah->attr.grh.dgid) : ah->attr.dlid;
        if (dlid == 0 || dlid >= SHIM_MAX_PEER) dlid = ah->attr.dlid;
        ud_send_one(dlid, buf, sizeof *h + len);
    }
    if ((wr->send_flags & IBV_SEND_SIGNALED) || ((shim_qp_t *)qp->shim)->init.sq_sig_all) complete(qp, wr, IBV_WC_SUCCESS, (uint32_t)len);
    return 0;
}

static const pub_mr_t *remote_mr(peer_t *p, uint32_t rkey, int refresh)
{
    uint32_t i = (rkey & 0xFFFF) - 1;
    if (i >= SHIM_MAX_MR) return NULL;
    pub_mr_t *c = &p->mr_cache[i % 64];
    if (p->pub_map) {                                               /* the live record: a load, never stale */
        *c = p->pub_map->mr[i];
        return __atomic_load_n(&p->pub_map->mr[i].rkey, __ATOMIC_ACQUIRE) == rkey && c->rkey == rkey ? c : NULL;
    }
    if (c->rkey == rkey && !refresh) return c;
    pub_mr_t m;
    if (peer_read(p, p->pub_addr + offsetof(pub_t, mr) + (uint64_t)i * sizeof m, &m, sizeof m)) return NULL;
    if (m.rkey != rkey) return NULL;
    *c = m;
    return c;
}

static int post_rc(struct ibv_qp *qp, struct ibv_send_wr *wr)
{
    shim_qp_t *q = qp->shim;
    if (q->attr.qp_state == IBV_QPS_ERR) { complete(qp, wr, IBV_WC_WR_FLUSH_ERR, 0); return 0; }
    if (q->attr.qp_state != IBV_QPS_RTS) return EINVAL;
    if (wr->opcode != IBV_WR_RDMA_WRITE && wr->opcode != IBV_WR_RDMA_READ) return ENOSYS;
    uint16_t dlid = lid_of_gid(&q->attr.ah_attr.grh.dgid);
    enum ibv_wc_status st = IBV_WC_SUCCESS;
    uint32_t total = 0;
    peer_t *p = peer_of(dlid, 0);
    for (int attempt = 0; attempt < 2; attempt++) {
        st = IBV_WC_SUCCESS; total = 0;
        if (!p) { st = IBV_WC_RETRY_EXC_ERR; break; }
        /* the responder's queue pair must be connected back to this one and willing to receive */
        pub_qp_t rq;
        uint32_t ridx = q->attr.dest_qp_num & 0xFFF;
        int unreachable = ridx >= SHIM_MAX_QP;
        if (!unreachable && p->pub_map) {
            rq = *(volatile pub_qp_t *)&p->pub_map->qp[ridx];
            if ((++p->ops & 63u) == 0 && kill(p->pid, 0) && errno == ESRCH) unreachable = 1;   /* a mapping outlives its owner */
        } else if (!unreachable) {
            unreachable = peer_read(p, p->pub_addr + offsetof(pub_t, qp) + (uint64_t)ridx * sizeof rq, &rq, sizeof rq) != 0;
        }
        if (unreachable) {
            st = IBV_WC_RETRY_EXC_ERR;
        } else if (rq.qpn != q->attr.dest_qp_num || rq.type != IBV_QPT_RC || (rq.state != IBV_QPS_RTR && rq.state != IBV_QPS_RTS) ||
                   rq.dest_qpn != qp->qp_num || rq.rq_psn != q->attr.sq_psn) {
            TRACE("qp %x -> %x refused: remote qpn %x state %u dest %x rq_psn %u (my sq_psn %u)\n", qp->qp_num, q->attr.dest_qp_num,
                  rq.qpn, rq.state, rq.dest_qpn, rq.rq_psn, q->attr.sq_psn);
            st = IBV_WC_RETRY_EXC_ERR;
            break;
        }
        if (st == IBV_WC_SUCCESS) {
            uint64_t raddr = wr->wr.rdma.remote_addr;
            for (int i = 0; i < wr->num_sge && st == IBV_WC_SUCCESS; i++) {
                uint32_t len = wr->sg_list[i].length;
                const pub_mr_t *m = remote_mr(p, wr->wr.rdma.rkey, attempt);
                uint32_t need = wr->opcode == IBV_WR_RDMA_WRITE ? IBV_ACCESS_REMOTE_WRITE : IBV_ACCESS_REMOTE_READ;
                if (!m || raddr < m->addr || raddr + len > m->addr + m->len || !(m->access & need)) { st = IBV_WC_REM_ACCESS_ERR; break; }
                uint8_t *seg = peer_seg(p, m);
                int rc = 0;
                if (seg) {
                    uint8_t *remote = seg + (raddr - m->shm_base), *local = (uint8_t *)(uintptr_t)wr->sg_list[i].addr;
                    if (wr->opcode == IBV_WR_RDMA_WRITE) memcpy(remote, local, len); else memcpy(local, remote, len);
                    __atomic_thread_fence(__ATOMIC_SEQ_CST);         /* placement order of consecutive work requests */
                    g_ops_memcpy++;
                } else {
                    g_ops_vm++;
                    rc = wr->opcode == IBV_WR_RDMA_WRITE ? peer_write(p, raddr, (void *)(uintptr_t)wr->sg_list[i].addr, len)
                                                         : peer_read(p, raddr, (void *)(uintptr_t)wr->sg_list[i].addr, len);
                }
                if (rc) st = (errno == ESRCH || errno == EPERM) ? IBV_WC_RETRY_EXC_ERR : IBV_WC_REM_ACCESS_ERR;
                raddr += len; total += len;
            }
        }
        if (st == IBV_WC_SUCCESS) break;
        if (attempt == 0) p = peer_of(dlid, 1);                      /* the peer may have restarted: re-read its record once */
    }
    if (st != IBV_WC_SUCCESS) {
        TRACE("qp %x op %d failed: %s\n", qp->qp_num, (int)wr->opcode, ibv_wc_status_str(st));
        complete(qp, wr, st, 0);                                     /* errors always complete */
        qp_to_error(qp);
        return 0;
    }
    if ((wr->send_flags & IBV_SEND_SIGNALED) || q->init.sq_sig_all) complete(qp, wr, IBV_WC_SUCCESS, total);
    return 0;
}

/* test hook (not a verbs call): how the RC operations of this process travelled so far */
void shim_transport_stats(uint64_t out[2]) { out[0] = g_ops_memcpy; out[1] = g_ops_vm; }

int ibv_post_send(struct ibv_qp *qp, struct ibv_send_wr *wr, struct ibv_send_wr **bad)
{
    for (; wr; wr = wr->next) {
        int rc = qp->qp_type == IBV_QPT_UD ? post_ud(qp, wr) : post_rc(qp, wr);
        if (rc) { if (bad) *bad = wr; return rc; }
    }
    return 0;
}

const char *ibv_wc_status_str(enum ibv_wc_status s)
{
    static const char *n[] = { "success", "local length error", "local QP operation error", "local EE context operation error",
        "local protection error", "Work Request Flushed Error", "memory management operation error", "bad response error",
        "local access error", "remote invalid request error", "remote access error", "remote operation error",
        "transport retry counter exceeded", "RNR retry counter exceeded", "local RDD violation error",
        "remote invalid RD request", "aborted error", "invalid EE context number", "invalid EE context state", "fatal error",
        "response timeout error", "general error" };
    return (unsigned)s < sizeof n / sizeof n[0] ? n[s] : "unknown";
}
