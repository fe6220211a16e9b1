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

/* ---- application-side driver (shared with the engine-side drop-in test: oracle/app_driver.inc) ---- */
#include "app_driver.inc"
int refstack_drive(void *proxy, int threads, int nconn, int64_t nreq, int plen, uint64_t *lat_ns, double *seconds)
{
    return app_drive(proxy, threads, nconn, nreq, plen, lat_ns, seconds, 3);
}
