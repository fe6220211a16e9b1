/*
 * oracle/ref_stack_access.c -- TEST INFRASTRUCTURE ONLY.
 * Read-only accessors into the RUNNING reference stack (oracle/_ref/libref_stack.so: the reference's unmodified
 * src/dare/*.c + proxy.c on the verbs shim): log offsets, log bytes, leadership.  Compiled against the reference's
 * own headers; the server state is reached the way the reference's transport code reaches it
 * (dare_ib_device->udata, dare_ibv_rc.c:42).
 */
#include <stdint.h>
#include <string.h>
#include <stdio.h>
#include <ev.h>
#include "dare_ibv.h"
#include "dare_server.h"
#include "dare_log.h"

extern dare_ib_device_t *dare_ib_device;
#define SRV ((dare_server_data_t *)dare_ib_device->udata)

int refstack_ready(void) { return dare_ib_device && dare_ib_device->udata && SRV->log && SRV->ctrl_data; }

/* head, apply, commit, end, tail, len, sid, term */
int refstack_offsets(uint64_t out[8])
{
    if (!refstack_ready()) return 1;
    dare_log_t *l = SRV->log;
    out[0] = l->head; out[1] = l->apply; out[2] = l->commit; out[3] = l->end; out[4] = l->tail; out[5] = l->len;
    out[6] = SRV->ctrl_data->sid; out[7] = SID_GET_TERM(SRV->ctrl_data->sid);
    return 0;
}

int refstack_log_read(uint64_t off, uint64_t n, void *dst)
{
    if (!refstack_ready() || off + n > SRV->log->len) return 1;
    memcpy(dst, SRV->log->entries + off, n);
    return 0;
}

int refstack_is_leader(void) { return refstack_ready() && is_leader(); }

int refstack_group_size(void) { return refstack_ready() ? (int)SRV->config.cid.size[0] : 0; }

/* ---- application-side driver: what src/spec_hooks.cpp does around an application's socket calls ------------------
 * `threads` application threads; connection c (fd 100+c) belongs to thread c % threads; request i goes to connection
 * i % nconn and carries payload bytes (i*31+k)&0xFF of length plen (plen >= 0) or of the ragged length
 * refstack_ragged_len(i, -plen).  Every proxy_on_read returns once the request is committed (proxy.c:160).
 * lat_ns[i] = latency of request i.  Returns 0. */
#include <pthread.h>
#include <stdlib.h>
#include <time.h>
struct proxy_node_t;
void proxy_on_read(struct proxy_node_t *proxy, void *buf, ssize_t bytes_read, int fd);
void proxy_on_accept(struct proxy_node_t *proxy, int fd);
void proxy_on_close(struct proxy_node_t *proxy, int fd);

uint32_t refstack_ragged_len(uint64_t i, uint32_t maxlen) { return (uint32_t)((i * 2654435761ull >> 7) % (maxlen + 1ull)); }

typedef struct { struct proxy_node_t *proxy; int t, threads, nconn; int64_t nreq; int plen; uint64_t *lat; } drv_t;
static uint64_t now_ns(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return (uint64_t)ts.tv_sec * 1000000000ull + ts.tv_nsec; }

static void *drv_thread(void *a)
{
    drv_t *d = (drv_t *)a;
    uint32_t cap = d->plen >= 0 ? (uint32_t)d->plen : (uint32_t)-d->plen;
    uint8_t *buf = (uint8_t *)malloc(cap + 1);
    for (int c = d->t; c < d->nconn; c += d->threads) proxy_on_accept(d->proxy, 100 + c);
    for (int64_t i = 0; i < d->nreq; i++) {
        int c = (int)(i % d->nconn);
        if (c % d->threads != d->t) continue;
        uint32_t len = d->plen >= 0 ? (uint32_t)d->plen : refstack_ragged_len((uint64_t)i, cap);
        if (len == 0) len = 1;                      /* a read() that returns 0 bytes is never forwarded (spec_hooks.cpp:168) */
        for (uint32_t k = 0; k < len; k++) buf[k] = (uint8_t)((i * 31 + k) & 0xFF);
        uint64_t t0 = now_ns();
        proxy_on_read(d->proxy, buf, (ssize_t)len, 100 + c);
        if (d->lat) d->lat[i] = now_ns() - t0;
    }
    for (int c = d->t; c < d->nconn; c += d->threads) proxy_on_close(d->proxy, 100 + c);
    free(buf);
    return NULL;
}

int refstack_drive(void *proxy, int threads, int nconn, int64_t nreq, int plen, uint64_t *lat_ns, double *seconds)
{
    if (threads < 1 || threads > 64 || nconn < threads) return 1;
    pthread_t th[64];
    drv_t d[64];
    uint64_t t0 = now_ns();
    for (int t = 0; t < threads; t++) {
        d[t] = (drv_t){ (struct proxy_node_t *)proxy, t, threads, nconn, nreq, plen, lat_ns };
        pthread_create(&th[t], NULL, drv_thread, &d[t]);
    }
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    if (seconds) *seconds = (double)(now_ns() - t0) * 1e-9;
    return 0;
}
