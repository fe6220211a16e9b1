/*
 * oracle/verbs_shim/shim_selftest.c -- TEST INFRASTRUCTURE ONLY: checks that the verbs stand-in behaves the way the
 * reference's transport code expects an HCA to behave (tests/test_verbs_shim.py builds and runs it).
 * Two processes ("ports" 1 and 2) connect a pair of RC queue pairs and a UD queue pair each:
 *   1  RDMA WRITE lands in the peer's registered memory; RDMA READ brings it back; signalled WRs complete, unsignalled don't
 *   2  a remote address outside the registered region completes with IBV_WC_REM_ACCESS_ERR, the QP enters ERR and the next
 *      WR is flushed (IBV_WC_WR_FLUSH_ERR)
 *   3  fencing: after the responder resets its QP (DARE's log-access revocation, dare_ibv_rc.c:2150) the requester's WRITE
 *      completes with IBV_WC_RETRY_EXC_ERR; the same for a PSN that does not match the responder's rq_psn
 *   4  UD: a unicast SEND arrives behind a 40-byte GRH with slid / src_qp filled in; a multicast SEND reaches every
 *      attached QP, the sender's own included
 * Prints "selftest ok" and exits 0.
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <sys/wait.h>
#include "infiniband/verbs.h"

#define CHECK(c) do { if (!(c)) { fprintf(stderr, "[port %d] FAILED line %d: %s\n", me, __LINE__, #c); exit(1); } } while (0)

typedef struct { uint32_t rc_qpn[3], ud_qpn; uint64_t addr; uint32_t rkey; uint16_t lid; union ibv_gid gid; } info_t;

static int me;
static int rd, wr;           /* pipe ends to the other process */
static void tell(const void *p, size_t n) { if (write(wr, p, n) != (ssize_t)n) exit(2); }
static void hear(void *p, size_t n) { size_t g = 0; while (g < n) { ssize_t k = read(rd, (char *)p + g, n - g); if (k <= 0) exit(2); g += (size_t)k; } }
static void barrier(void) { char c = 'x'; tell(&c, 1); hear(&c, 1); }

static int poll1(struct ibv_cq *cq, struct ibv_wc *wc)
{
    for (int i = 0; i < 2000000; i++) { int n = ibv_poll_cq(cq, 1, wc); if (n) return n; }
    return 0;
}

static void connect_qp(struct ibv_qp *qp, const info_t *peer, uint32_t dest_qpn, uint32_t rq_psn, uint32_t sq_psn)
{
    struct ibv_qp_attr a;
    memset(&a, 0, sizeof a);
    a.qp_state = IBV_QPS_RESET;
    CHECK(ibv_modify_qp(qp, &a, IBV_QP_STATE) == 0);
    a.qp_state = IBV_QPS_INIT; a.port_num = 1;
    a.qp_access_flags = IBV_ACCESS_REMOTE_WRITE | IBV_ACCESS_REMOTE_READ | IBV_ACCESS_LOCAL_WRITE;
    CHECK(ibv_modify_qp(qp, &a, IBV_QP_STATE | IBV_QP_PKEY_INDEX | IBV_QP_PORT | IBV_QP_ACCESS_FLAGS) == 0);
    memset(&a, 0, sizeof a);
    a.qp_state = IBV_QPS_RTR; a.path_mtu = IBV_MTU_4096; a.dest_qp_num = dest_qpn; a.rq_psn = rq_psn;
    a.ah_attr.is_global = 1; a.ah_attr.grh.dgid = peer->gid; a.ah_attr.dlid = peer->lid; a.ah_attr.port_num = 1;
    CHECK(ibv_modify_qp(qp, &a, IBV_QP_STATE | IBV_QP_PATH_MTU | IBV_QP_MAX_DEST_RD_ATOMIC | IBV_QP_MIN_RNR_TIMER |
                        IBV_QP_RQ_PSN | IBV_QP_AV | IBV_QP_DEST_QPN) == 0);
    memset(&a, 0, sizeof a);
    a.qp_state = IBV_QPS_RTS; a.sq_psn = sq_psn;
    CHECK(ibv_modify_qp(qp, &a, IBV_QP_STATE | IBV_QP_TIMEOUT | IBV_QP_RETRY_CNT | IBV_QP_RNR_RETRY | IBV_QP_SQ_PSN |
                        IBV_QP_MAX_QP_RD_ATOMIC) == 0);
}

static int rdma(struct ibv_qp *qp, struct ibv_mr *mr, void *buf, uint32_t len, enum ibv_wr_opcode op, uint64_t raddr, uint32_t rkey,
                int signaled)
{
    struct ibv_sge sg = { (uint64_t)(uintptr_t)buf, len, mr->lkey };
    struct ibv_send_wr w, *bad;
    memset(&w, 0, sizeof w);
    w.wr_id = 77; w.sg_list = &sg; w.num_sge = 1; w.opcode = op; w.send_flags = signaled ? IBV_SEND_SIGNALED : 0;
    w.wr.rdma.remote_addr = raddr; w.wr.rdma.rkey = rkey;
    return ibv_post_send(qp, &w, &bad);
}

int main(void)
{
    int ab[2], ba[2];
    if (pipe(ab) || pipe(ba)) return 2;
    pid_t child = fork();
    me = child ? 1 : 2;
    rd = child ? ba[0] : ab[0]; wr = child ? ab[1] : ba[1];
    char id[8]; snprintf(id, sizeof id, "%d", me); setenv("APUS_SHIM_ID", id, 1);

    int ndev = 0;
    struct ibv_device **dl = ibv_get_device_list(&ndev);
    CHECK(ndev == 1);
    struct ibv_context *ctx = ibv_open_device(dl[0]);
    struct ibv_port_attr pa; CHECK(ibv_query_port(ctx, 1, &pa) == 0 && pa.lid == me && pa.state == IBV_PORT_ACTIVE);
    struct ibv_pd *pd = ibv_alloc_pd(ctx);
    static uint8_t region[8192] __attribute__((aligned(4096)));   /* whole pages of its own: the shm transport re-backs them */
    static uint8_t local[4096];
    struct ibv_mr *mr = ibv_reg_mr(pd, region, sizeof region, IBV_ACCESS_REMOTE_WRITE | IBV_ACCESS_REMOTE_READ | IBV_ACCESS_LOCAL_WRITE);
    struct ibv_mr *lmr = ibv_reg_mr(pd, local, sizeof local, IBV_ACCESS_LOCAL_WRITE);
    struct ibv_cq *cq = ibv_create_cq(ctx, 64, NULL, NULL, 0), *ucq_s = ibv_create_cq(ctx, 64, NULL, NULL, 0),
                  *ucq_r = ibv_create_cq(ctx, 64, NULL, NULL, 0);
    struct ibv_qp_init_attr ia;
    memset(&ia, 0, sizeof ia);
    ia.qp_type = IBV_QPT_RC; ia.send_cq = ia.recv_cq = cq; ia.cap.max_send_wr = 16; ia.cap.max_recv_wr = 1; ia.cap.max_send_sge = ia.cap.max_recv_sge = 1;
    ia.cap.max_inline_data = 1 << 20;
    CHECK(ibv_create_qp(pd, &ia) == NULL);                        /* find_max_inline() expects oversize requests to fail */
    ia.cap.max_inline_data = 256;
    struct ibv_qp *qp[3];
    for (int i = 0; i < 3; i++) CHECK((qp[i] = ibv_create_qp(pd, &ia)) != NULL);
    memset(&ia, 0, sizeof ia);
    ia.qp_type = IBV_QPT_UD; ia.send_cq = ucq_s; ia.recv_cq = ucq_r; ia.cap.max_send_wr = ia.cap.max_recv_wr = 16; ia.cap.max_send_sge = ia.cap.max_recv_sge = 1;
    struct ibv_qp *ud = ibv_create_qp(pd, &ia);
    CHECK(ud);
    union ibv_gid mg; memset(&mg, 0, sizeof mg); mg.raw[0] = 0xff; mg.raw[1] = 0x0e;
    CHECK(ibv_attach_mcast(ud, &mg, 0xc001) == 0);
    struct ibv_qp_attr a; memset(&a, 0, sizeof a);
    a.qp_state = IBV_QPS_INIT; a.port_num = 1; CHECK(ibv_modify_qp(ud, &a, IBV_QP_STATE | IBV_QP_PKEY_INDEX | IBV_QP_PORT | IBV_QP_QKEY) == 0);
    a.qp_state = IBV_QPS_RTR; CHECK(ibv_modify_qp(ud, &a, IBV_QP_STATE) == 0);
    a.qp_state = IBV_QPS_RTS; CHECK(ibv_modify_qp(ud, &a, IBV_QP_STATE | IBV_QP_SQ_PSN) == 0);
    static uint8_t rbuf[4][512];
    for (int i = 0; i < 4; i++) {
        struct ibv_sge sg = { (uint64_t)(uintptr_t)rbuf[i], sizeof rbuf[i], lmr->lkey };
        struct ibv_recv_wr rw = { (uint64_t)i, NULL, &sg, 1 }, *bad;
        CHECK(ibv_post_recv(ud, &rw, &bad) == 0);
    }

    info_t mine, peer;
    memset(&mine, 0, sizeof mine);
    for (int i = 0; i < 3; i++) mine.rc_qpn[i] = qp[i]->qp_num;
    mine.ud_qpn = ud->qp_num; mine.addr = (uint64_t)(uintptr_t)region; mine.rkey = mr->rkey; mine.lid = pa.lid;
    CHECK(ibv_query_gid(ctx, 1, 0, &mine.gid) == 0);
    tell(&mine, sizeof mine); hear(&peer, sizeof peer);
    connect_qp(qp[0], &peer, peer.rc_qpn[0], 100, 100);
    connect_qp(qp[1], &peer, peer.rc_qpn[1], 200, 200);
    /* pair 2: port 1 will send with PSN 5 while port 2 expects 6 */
    connect_qp(qp[2], &peer, peer.rc_qpn[2], me == 2 ? 6 : 5, 5);
    barrier();

    struct ibv_wc wc;
    if (me == 1) {
        /* 1: WRITE (unsignalled, then signalled), READ back */
        for (int i = 0; i < 64; i++) local[i] = (uint8_t)(i * 3 + 1);
        CHECK(rdma(qp[0], lmr, local, 64, IBV_WR_RDMA_WRITE, peer.addr + 128, peer.rkey, 0) == 0);
        CHECK(ibv_poll_cq(cq, 1, &wc) == 0);                                     /* unsignalled success: no CQE */
        CHECK(rdma(qp[0], lmr, local, 64, IBV_WR_RDMA_WRITE, peer.addr + 256, peer.rkey, 1) == 0);
        CHECK(poll1(cq, &wc) == 1 && wc.status == IBV_WC_SUCCESS && wc.wr_id == 77 && wc.opcode == IBV_WC_RDMA_WRITE);
        memset(local + 1024, 0, 64);
        CHECK(rdma(qp[0], lmr, local + 1024, 64, IBV_WR_RDMA_READ, peer.addr + 128, peer.rkey, 1) == 0);
        CHECK(poll1(cq, &wc) == 1 && wc.status == IBV_WC_SUCCESS && wc.opcode == IBV_WC_RDMA_READ);
        CHECK(memcmp(local, local + 1024, 64) == 0);
        barrier();                                                                   /* A */
        /* 2: out of the registered range, wrong rkey */
        CHECK(rdma(qp[0], lmr, local, 64, IBV_WR_RDMA_WRITE, peer.addr + 8186, peer.rkey, 0) == 0);
        CHECK(poll1(cq, &wc) == 1 && wc.status == IBV_WC_REM_ACCESS_ERR);         /* errors complete even when unsignalled */
        struct ibv_qp_attr qa; struct ibv_qp_init_attr qi;
        CHECK(ibv_query_qp(qp[0], &qa, IBV_QP_STATE, &qi) == 0 && qa.qp_state == IBV_QPS_ERR);
        CHECK(rdma(qp[0], lmr, local, 64, IBV_WR_RDMA_WRITE, peer.addr, peer.rkey, 0) == 0);
        CHECK(poll1(cq, &wc) == 1 && wc.status == IBV_WC_WR_FLUSH_ERR);
        /* 3: the responder revokes access to pair 1, then pair 2 has the wrong PSN */
        CHECK(rdma(qp[1], lmr, local, 64, IBV_WR_RDMA_WRITE, peer.addr, peer.rkey, 1) == 0);
        CHECK(poll1(cq, &wc) == 1 && wc.status == IBV_WC_SUCCESS);
        barrier();                                                                   /* B: peer resets its qp[1] */
        barrier();                                                                   /* C */
        CHECK(rdma(qp[1], lmr, local, 64, IBV_WR_RDMA_WRITE, peer.addr, peer.rkey, 1) == 0);
        CHECK(poll1(cq, &wc) == 1 && wc.status == IBV_WC_RETRY_EXC_ERR);
        CHECK(rdma(qp[2], lmr, local, 64, IBV_WR_RDMA_WRITE, peer.addr, peer.rkey, 1) == 0);
        CHECK(poll1(cq, &wc) == 1 && wc.status == IBV_WC_RETRY_EXC_ERR);
        /* a QP that was never connected cannot post */
        memset(&a, 0, sizeof a); a.qp_state = IBV_QPS_RESET; CHECK(ibv_modify_qp(qp[2], &a, IBV_QP_STATE) == 0);
        CHECK(rdma(qp[2], lmr, local, 64, IBV_WR_RDMA_WRITE, peer.addr, peer.rkey, 1) != 0);
        /* 4: UD unicast then multicast */
        struct ibv_ah_attr aa; memset(&aa, 0, sizeof aa);
        aa.is_global = 1; aa.dlid = peer.lid; aa.grh.dgid = peer.gid; aa.port_num = 1;
        struct ibv_ah *ah = ibv_create_ah(pd, &aa);
        memcpy(local, "unicast-hello", 14);
        struct ibv_sge sg = { (uint64_t)(uintptr_t)local, 14, lmr->lkey };
        struct ibv_send_wr w, *bad; memset(&w, 0, sizeof w);
        w.wr_id = 5; w.sg_list = &sg; w.num_sge = 1; w.opcode = IBV_WR_SEND; w.send_flags = IBV_SEND_SIGNALED;
        w.wr.ud.ah = ah; w.wr.ud.remote_qpn = peer.ud_qpn;
        CHECK(ibv_post_send(ud, &w, &bad) == 0);
        CHECK(poll1(ucq_s, &wc) == 1 && wc.status == IBV_WC_SUCCESS && wc.wr_id == 5);
        memset(&aa, 0, sizeof aa); aa.is_global = 1; aa.dlid = 0xc001; aa.grh.dgid = mg; aa.port_num = 1;
        struct ibv_ah *mah = ibv_create_ah(pd, &aa);
        memcpy(local, "multicast-hello", 16); sg.length = 16;
        w.wr.ud.ah = mah; w.wr.ud.remote_qpn = 0xFFFFFF;
        CHECK(ibv_post_send(ud, &w, &bad) == 0);
        CHECK(poll1(ucq_s, &wc) == 1 && wc.status == IBV_WC_SUCCESS);
        CHECK(poll1(ucq_r, &wc) == 1 && wc.status == IBV_WC_SUCCESS && wc.slid == 1 && wc.byte_len == 16 + 40);   /* my own multicast */
        CHECK(memcmp(rbuf[wc.wr_id] + 40, "multicast-hello", 16) == 0);
        barrier();                                                                   /* D */
    } else {
        barrier();                                                                   /* A */
        for (int i = 0; i < 64; i++) CHECK(region[128 + i] == (uint8_t)(i * 3 + 1) && region[256 + i] == (uint8_t)(i * 3 + 1));
        barrier();                                                                   /* B */
        memset(&a, 0, sizeof a); a.qp_state = IBV_QPS_RESET;
        CHECK(ibv_modify_qp(qp[1], &a, IBV_QP_STATE) == 0);                        /* revoke */
        barrier();                                                                   /* C */
        CHECK(poll1(ucq_r, &wc) == 1 && wc.status == IBV_WC_SUCCESS && (wc.opcode & IBV_WC_RECV) && wc.slid == 1 &&
              wc.src_qp == peer.ud_qpn && wc.byte_len == 14 + 40);
        CHECK(memcmp(rbuf[wc.wr_id] + 40, "unicast-hello", 14) == 0);
        struct ibv_grh *g = (struct ibv_grh *)rbuf[wc.wr_id];
        CHECK(memcmp(g->sgid.raw, peer.gid.raw, 16) == 0);
        CHECK(poll1(ucq_r, &wc) == 1 && wc.status == IBV_WC_SUCCESS && wc.slid == 1 && wc.byte_len == 16 + 40);
        CHECK(memcmp(rbuf[wc.wr_id] + 40, "multicast-hello", 16) == 0);
        barrier();                                                                   /* D */
    }
    ibv_dereg_mr(mr); ibv_dereg_mr(lmr);
    if (me == 1) {
        int st = 0;
        waitpid(child, &st, 0);
        ibv_close_device(ctx);
        if (!WIFEXITED(st) || WEXITSTATUS(st)) { fprintf(stderr, "port 2 failed\n"); return 1; }
        printf("selftest ok\n");
    } else {
        ibv_close_device(ctx);
    }
    return 0;
}
