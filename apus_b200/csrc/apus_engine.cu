/*
 * apus_engine.cu -- host side of the C ABI declared in include/apus_gpu.h.
 *
 * Thin by design: allocate the replica's HBM region, map peers (peer access or
 * CUDA IPC -- the replacement for ibv_reg_mr + the raddr/rkey exchange of
 * dare_ibv_rc.c:245-273 / dare_ibv_ud.c:1116-1119), feed the leader's submission
 * ring (the replacement for the malloc'd tailq_entry_t list of message.h:11-22)
 * and launch / stop / observe the persistent kernels of apus_kernels.cu.
 * All replication work happens in those kernels; nothing here touches entry
 * bytes, and there is no CPU fallback.
 */
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "apus_gpu.h"
#include "apus_layout.h"

extern "C" cudaError_t apus_launch_roles(const apus_role_t *d_roles, int n_roles, cudaStream_t stream);
extern "C" size_t apus_kernel_smem_bytes(void);

static __thread char g_err[512];

static int fail(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return APUS_ERROR;
}

#define CK(call)                                                                             \
    do {                                                                                     \
        cudaError_t _e = (call);                                                             \
        if (_e != cudaSuccess)                                                               \
            return fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

struct DeviceGuard {
    int prev;
    bool ok;
    explicit DeviceGuard(int dev) : prev(-1), ok(false)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = (cudaSetDevice(dev) == cudaSuccess);
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

#define PEER_MAGIC 0x4150555342323030ull /* "APUSB200" */

struct peer_blob {
    uint64_t magic;
    int64_t  pid;
    int32_t  device;
    uint32_t pad;
    uint64_t ptr;
    uint64_t bytes;
    cudaIpcMemHandle_t ipc;
};
static_assert(sizeof(peer_blob) <= sizeof(apus_peer_handle_t), "peer handle too small");

struct apus_replica {
    apus_config_t cfg;
    uint64_t log_len;
    uint64_t entries_off;
    uint32_t idx_cap;
    uint8_t *region;
    size_t   region_bytes;
    apus_devctx_t  h_ctx;
    apus_devctx_t *d_ctx;
    apus_role_t   *d_roles;       /* table used when this replica owns a launch */
    apus_hostwords_t *hw;         /* pinned + mapped */
    apus_hostwords_t *hw_dev;
    /* submission ring (leader) */
    apus_slot_t *ring_desc_host;  /* pinned: the slot ring itself (mapped mode) or its staging mirror */
    uint8_t     *ring_pay_host;
    apus_slot_t *ring_desc_dev;   /* device-visible address the kernel reads */
    uint8_t     *ring_pay_dev;
    uint64_t    *sub_tail_dev;    /* device doorbell (device mode) */
    uint64_t    *sub_tail_stage;  /* pinned staging word for the device doorbell */
    uint32_t ring_slots, ring_bytes;
    uint64_t submitted, flushed;
    uint64_t pay_head;            /* payload bytes handed out (monotone; position = % ring_bytes) */
    uint64_t pay_flushed;         /* payload bytes already made visible to the kernel */
    uint64_t *pay_end;            /* [ticket & mask] = pay_head after that ticket's image */
    int      defer;
    /* launch */
    cudaStream_t stream, copy_stream;
    cudaEvent_t  ev_start, ev_stop;
    apus_replica *launch_owner;   /* replica whose stream/events carry the launch */
    int      in_flight;
    uint64_t launches;
    void    *peer_ptr[APUS_MAX_SERVERS];
    int      peer_is_ipc[APUS_MAX_SERVERS];
    uint32_t *d_lat;
};

extern "C" int apus_abi_version(void) { return APUS_ABI_VERSION; }
extern "C" const char *apus_last_error(void) { return g_err; }
extern "C" int apus_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

static inline int is_leader(const apus_replica *r) { return r->cfg.server_idx == r->cfg.leader_idx; }

extern "C" int apus_replica_create(const apus_config_t *cfg, apus_replica_t **out)
{
    if (!cfg || !out) return fail("null argument");
    if (cfg->struct_size != sizeof(apus_config_t)) return fail("apus_config_t size mismatch (ABI)");
    if (cfg->group_size < 1 || cfg->group_size > APUS_MAX_SERVER_COUNT) return fail("group_size out of range");
    if (cfg->server_idx >= cfg->group_size || cfg->leader_idx >= cfg->group_size) return fail("bad server/leader idx");
    int ndev = apus_device_count();
    if (ndev <= 0) return fail("no CUDA device: the engine has no CPU fallback");
    if (cfg->device < 0 || cfg->device >= ndev) return fail("device %d out of range (%d present)", cfg->device, ndev);
    uint64_t log_len = cfg->log_size ? cfg->log_size : APUS_LOG_SIZE;
    if (log_len % 4096 || log_len < 8192) return fail("log_size must be a multiple of 4096 (>= 8192)");

    DeviceGuard g(cfg->device);
    if (!g.ok) return fail("cudaSetDevice(%d) failed", cfg->device);
    apus_replica *r = (apus_replica *)calloc(1, sizeof(*r));
    if (!r) return fail("out of memory");
    r->cfg = *cfg;
    if (!(r->cfg.flags & APUS_F_EXPLICIT)) r->cfg.flags |= APUS_F_DEVICE_STATS;
    r->log_len = log_len;
    if (log_len > (1ull << 31)) return fail("log_size above 2 GiB is not supported (32-bit offset index)");
    uint32_t cap = 1024;
    while ((uint64_t)cap * 64ull < log_len) cap <<= 1;
    r->idx_cap = cap;
    r->entries_off = (APUS_INDEX_OFF + (uint64_t)cap * 4ull + 4095ull) & ~4095ull;
    r->region_bytes = r->entries_off + log_len;
    CK(cudaMalloc(&r->region, r->region_bytes));
    CK(cudaMemset(r->region, 0, r->region_bytes));
    /* log_new(): end = tail = old_end = len (dare_log.h:129-134) */
    apus_loghdr_t h;
    memset(&h, 0, sizeof h);
    h.len = log_len; h.end = log_len; h.tail = log_len; h.old_end = log_len;
    CK(cudaMemcpy(r->region + APUS_HDR_OFF, &h, sizeof h, cudaMemcpyHostToDevice));
    apus_ctrl_t c;
    memset(&c, 0, sizeof c);
    c.next_idx = 1;
    c.pend_head_end = log_len;   /* no HEAD entry pending */
    CK(cudaMemcpy(r->region, &c, sizeof c, cudaMemcpyHostToDevice));

    CK(cudaHostAlloc(&r->hw, sizeof(apus_hostwords_t), cudaHostAllocMapped | cudaHostAllocPortable));
    memset((void *)r->hw, 0, sizeof(apus_hostwords_t));
    CK(cudaHostGetDevicePointer(&r->hw_dev, r->hw, 0));
    CK(cudaMalloc(&r->d_ctx, sizeof(apus_devctx_t)));
    CK(cudaMalloc(&r->d_roles, sizeof(apus_role_t) * 64));
    CK(cudaMalloc(&r->d_lat, sizeof(uint32_t) * APUS_LAT_RING));
    CK(cudaMemset(r->d_lat, 0, sizeof(uint32_t) * APUS_LAT_RING));
    CK(cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&r->copy_stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&r->ev_start));
    CK(cudaEventCreate(&r->ev_stop));

    if (is_leader(r)) {
        uint32_t slots = cfg->ring_slots ? cfg->ring_slots : (1u << 16);
        uint32_t bytes = cfg->ring_bytes ? cfg->ring_bytes : (16u << 20);
        if (slots & (slots - 1)) return fail("ring_slots must be a power of two");
        if (bytes % 4096 || bytes < (1u << 17)) return fail("ring_bytes must be a multiple of 4096, >= 128 KiB");
        if (bytes / 16 > 0x00ffffffu) return fail("ring_bytes too large for the 24-bit descriptor offset");
        r->ring_slots = slots; r->ring_bytes = bytes;
        r->pay_end = (uint64_t *)calloc(slots, sizeof(uint64_t));
        if (!r->pay_end) return fail("out of memory");
        CK(cudaHostAlloc(&r->ring_desc_host, sizeof(apus_slot_t) * slots, cudaHostAllocMapped | cudaHostAllocPortable));
        CK(cudaHostAlloc(&r->ring_pay_host, bytes, cudaHostAllocMapped | cudaHostAllocPortable));
        if (cfg->ring_mode == APUS_RING_HOST_MAPPED) {
            CK(cudaHostGetDevicePointer(&r->ring_desc_dev, r->ring_desc_host, 0));
            CK(cudaHostGetDevicePointer(&r->ring_pay_dev, r->ring_pay_host, 0));
        } else {
            CK(cudaMalloc(&r->ring_desc_dev, sizeof(apus_slot_t) * slots));
            CK(cudaMalloc(&r->ring_pay_dev, bytes));
            CK(cudaMalloc(&r->sub_tail_dev, 128));
            CK(cudaMemset(r->sub_tail_dev, 0, 128));
            CK(cudaHostAlloc(&r->sub_tail_stage, 64, cudaHostAllocPortable));
        }
    }
    r->peer_ptr[cfg->server_idx] = r->region;
    *out = r;
    return APUS_OK;
}

extern "C" void apus_replica_destroy(apus_replica_t *r)
{
    if (!r) return;
    DeviceGuard g(r->cfg.device);
    if (r->in_flight) {
        r->hw->stop = 1;
        cudaEventSynchronize(r->launch_owner ? r->launch_owner->ev_stop : r->ev_stop);
    }
    for (int i = 0; i < APUS_MAX_SERVER_COUNT; i++)
        if (r->peer_is_ipc[i] && r->peer_ptr[i]) cudaIpcCloseMemHandle(r->peer_ptr[i]);
    if (r->cfg.ring_mode != APUS_RING_HOST_MAPPED) {
        cudaFree(r->ring_desc_dev); cudaFree(r->ring_pay_dev); cudaFree(r->sub_tail_dev);
        if (r->sub_tail_stage) cudaFreeHost(r->sub_tail_stage);
    }
    if (r->ring_desc_host) cudaFreeHost(r->ring_desc_host);
    if (r->ring_pay_host) cudaFreeHost(r->ring_pay_host);
    cudaFree(r->d_lat); cudaFree(r->d_roles); cudaFree(r->d_ctx);
    cudaEventDestroy(r->ev_start); cudaEventDestroy(r->ev_stop);
    cudaStreamDestroy(r->stream); cudaStreamDestroy(r->copy_stream);
    cudaFreeHost((void *)r->hw);
    cudaFree(r->region);
    free(r->pay_end);
    free(r);
}

extern "C" int apus_replica_export(apus_replica_t *r, apus_peer_handle_t *out)
{
    if (!r || !out) return fail("null argument");
    DeviceGuard g(r->cfg.device);
    peer_blob b;
    memset(&b, 0, sizeof b);
    b.magic = PEER_MAGIC; b.pid = (int64_t)getpid(); b.device = r->cfg.device;
    b.ptr = (uint64_t)(uintptr_t)r->region; b.bytes = r->region_bytes;
    CK(cudaIpcGetMemHandle(&b.ipc, r->region));
    memset(out, 0, sizeof *out);
    memcpy(out, &b, sizeof b);
    return APUS_OK;
}

extern "C" int apus_replica_connect(apus_replica_t *r, uint8_t peer_idx, const apus_peer_handle_t *peer)
{
    if (!r || !peer) return fail("null argument");
    if (peer_idx >= r->cfg.group_size) return fail("peer idx out of range");
    if (peer_idx == r->cfg.server_idx) return APUS_OK;
    peer_blob b;
    memcpy(&b, peer, sizeof b);
    if (b.magic != PEER_MAGIC) return fail("bad peer handle");
    if (b.bytes != r->region_bytes) return fail("peer region size differs (log_size mismatch)");
    DeviceGuard g(r->cfg.device);
    if (b.pid == (int64_t)getpid()) {
        if (b.device != r->cfg.device) {
            int can = 0;
            CK(cudaDeviceCanAccessPeer(&can, r->cfg.device, b.device));
            if (!can) return fail("device %d cannot access peer device %d", r->cfg.device, b.device);
            cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                return fail("cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
            cudaGetLastError();
        }
        r->peer_ptr[peer_idx] = (void *)(uintptr_t)b.ptr;
        r->peer_is_ipc[peer_idx] = 0;
    } else {
        void *p = NULL;
        CK(cudaIpcOpenMemHandle(&p, b.ipc, cudaIpcMemLazyEnablePeerAccess));
        r->peer_ptr[peer_idx] = p;
        r->peer_is_ipc[peer_idx] = 1;
    }
    return APUS_OK;
}

static void fill_ctx(apus_replica *r, uint64_t target)
{
    apus_devctx_t *c = &r->h_ctx;
    memset(c, 0, sizeof *c);
    c->idx = r->cfg.server_idx; c->group_size = r->cfg.group_size; c->leader_idx = r->cfg.leader_idx;
    c->quorum = (uint8_t)(r->cfg.group_size / 2 + 1);        /* dare_ibv_rc.c:1741 */
    c->flags = r->cfg.flags & 0x7fffffffu;
    c->term = r->cfg.term; c->log_len = r->log_len; c->target = target;
    c->entries_off = r->entries_off; c->idx_mask = r->idx_cap - 1;
    c->n_workers = r->cfg.leader_ctas ? r->cfg.leader_ctas : 4;
    if (c->n_workers > 32) c->n_workers = 32;
    c->epoch = (uint32_t)(r->launches + 1);
    c->doorbell_relay = (r->cfg.ring_mode == APUS_RING_HOST_MAPPED && c->n_workers >= 2) ? 1u : 0u;
    c->region = r->region;
    for (int i = 0; i < APUS_MAX_SERVER_COUNT; i++)
        c->peer[i] = (i == r->cfg.server_idx) ? NULL : (uint8_t *)r->peer_ptr[i];
    c->sub_slots = r->ring_desc_dev; c->sub_pay = r->ring_pay_dev;
    c->sub_mask = r->ring_slots ? r->ring_slots - 1 : 0;
    c->sub_tail = (r->cfg.ring_mode == APUS_RING_HOST_MAPPED) ? (const volatile uint64_t *)&r->hw_dev->sub_tail
                                                              : (const volatile uint64_t *)r->sub_tail_dev;
    c->hw = r->hw_dev;
    c->lat_ns = r->d_lat;
}

extern "C" int apus_replicas_launch(apus_replica_t **rs, int n, uint64_t target)
{
    if (!rs || n < 1 || n > 64) return fail("bad replica list");
    apus_replica *owner = rs[0];
    for (int i = 0; i < n; i++) {
        if (!rs[i]) return fail("null replica");
        if (rs[i]->cfg.device != owner->cfg.device) return fail("replicas of one launch must share a device");
        if (rs[i]->in_flight) return fail("replica %d already has a launch in flight", (int)rs[i]->cfg.server_idx);
        if (!is_leader(rs[i]) && !rs[i]->peer_ptr[rs[i]->cfg.leader_idx]) return fail("follower not connected to its leader");
    }
    DeviceGuard g(owner->cfg.device);
    apus_role_t roles[64];
    int nroles = 0;
    for (int i = 0; i < n; i++) {
        apus_replica *r = rs[i];
        fill_ctx(r, target);
        r->hw->stop = 0; r->hw->error = 0;
        CK(cudaMemcpyAsync(r->d_ctx, &r->h_ctx, sizeof(apus_devctx_t), cudaMemcpyHostToDevice, owner->stream));
        if (is_leader(r)) {
            for (uint32_t w = 0; w < r->h_ctx.n_workers; w++) {
                if (nroles >= 64) return fail("too many roles in one launch");
                roles[nroles].kind = APUS_ROLE_LEADER; roles[nroles].worker = w; roles[nroles].ctx = r->d_ctx; nroles++;
            }
        } else {
            if (nroles >= 64) return fail("too many roles in one launch");
            roles[nroles].kind = APUS_ROLE_FOLLOWER; roles[nroles].worker = 0; roles[nroles].ctx = r->d_ctx; nroles++;
        }
    }
    CK(cudaMemcpyAsync(owner->d_roles, roles, sizeof(apus_role_t) * nroles, cudaMemcpyHostToDevice, owner->stream));
    /* roles[] is on the stack: the copy above must have read it before we return */
    CK(cudaStreamSynchronize(owner->stream));
    CK(cudaEventRecord(owner->ev_start, owner->stream));
    CK(apus_launch_roles(owner->d_roles, nroles, owner->stream));
    CK(cudaEventRecord(owner->ev_stop, owner->stream));
    for (int i = 0; i < n; i++) {
        rs[i]->launch_owner = owner;
        rs[i]->in_flight = 1;
        rs[i]->launches++;
    }
    return APUS_OK;
}

extern "C" int apus_replica_wait(apus_replica_t *r, int64_t timeout_ms)
{
    if (!r) return fail("null argument");
    if (!r->in_flight) return APUS_OK;
    apus_replica *o = r->launch_owner;
    DeviceGuard g(o->cfg.device);
    if (timeout_ms < 0) {
        CK(cudaEventSynchronize(o->ev_stop));
    } else {
        struct timespec t0, t;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        for (;;) {
            cudaError_t e = cudaEventQuery(o->ev_stop);
            if (e == cudaSuccess) break;
            if (e != cudaErrorNotReady) return fail("kernel failed: %s", cudaGetErrorString(e));
            clock_gettime(CLOCK_MONOTONIC, &t);
            int64_t ms = (t.tv_sec - t0.tv_sec) * 1000 + (t.tv_nsec - t0.tv_nsec) / 1000000;
            if (ms > timeout_ms) { snprintf(g_err, sizeof g_err, "timeout"); return APUS_RETRY; }
            usleep(50);
        }
    }
    r->in_flight = 0;
    if (r->hw->error) return fail("kernel reported protocol error %llu", (unsigned long long)r->hw->error);
    return APUS_OK;
}

extern "C" int apus_replica_last_launch_ms(apus_replica_t *r, float *ms)
{
    if (!r || !ms || !r->launch_owner) return fail("no launch");
    DeviceGuard g(r->launch_owner->cfg.device);
    CK(cudaEventElapsedTime(ms, r->launch_owner->ev_start, r->launch_owner->ev_stop));
    return APUS_OK;
}

extern "C" int apus_replicas_stop(apus_replica_t **rs, int n)
{
    if (!rs) return fail("null argument");
    for (int i = 0; i < n; i++) rs[i]->hw->stop = 1;
    __sync_synchronize();
    int rc = APUS_OK;
    for (int i = 0; i < n; i++) {
        int e = apus_replica_wait(rs[i], 30000);
        if (e != APUS_OK) rc = e;
    }
    return rc;
}

/* ---- submission ------------------------------------------------------------------ */
static inline uint32_t image_bytes(uint8_t type, uint16_t len)
{
    if (type == APUS_NOOP) return 0;
    if (type == APUS_CONFIG) return 16;
    if (type == APUS_HEAD) return 8;
    return 2u + len;
}

/* Requests whose data image fits APUS_SLOT_INLINE travel inside their 128 B slot;
 * larger images go to the payload byte ring.  Payload space is tracked with monotone
 * byte counters: pay_head (bytes handed out, skip gaps included) and, per ticket, the
 * counter value after its image. */
static int ring_put(apus_replica *r, uint8_t type, uint16_t conn, uint64_t req_id, const void *cmd, uint16_t len)
{
    const uint64_t consumed = r->hw->consumed;
    const uint32_t mask = r->ring_slots - 1;
    if (r->submitted - consumed >= r->ring_slots) return APUS_RETRY;
    const uint32_t nb = image_bytes(type, len);
    apus_slot_t *d = &r->ring_desc_host[r->submitted & mask];
    uint32_t type_off = ((uint32_t)type & APUS_SLOT_TYPE_MASK) << APUS_SLOT_TYPE_SHIFT;
    uint64_t head = r->pay_head;
    uint8_t *dst = d->inl;
    if (nb > APUS_SLOT_INLINE) {
        const uint32_t need = (nb + 15u) & ~15u;
        const uint64_t R = r->ring_bytes;
        uint64_t pos = head % R;
        const uint64_t skip = (pos + need > R) ? (R - pos) : 0;          /* an image never wraps */
        const uint64_t tail = consumed ? r->pay_end[(consumed - 1) & mask] : 0;
        if ((head - tail) + skip + need > R) return APUS_RETRY;
        head += skip;
        pos = head % R;
        dst = r->ring_pay_host + pos;
        type_off |= APUS_SLOT_EXT | (uint32_t)(pos / 16);
        if (skip || (pos == 0 && head != 0)) type_off |= APUS_SLOT_WRAP;
        head += need;
    }
    if (nb) {
        if (type == APUS_CONFIG || type == APUS_HEAD) {
            memcpy(dst, cmd, nb);
        } else {
            memcpy(dst, &len, 2);                    /* sm_cmd_t {u16 len; u8 cmd[]} (dare_sm.h:23-27) */
            if (len) memcpy(dst + 2, cmd, len);
        }
    }
    d->req_id = req_id;
    d->type_off = type_off;
    d->len = len;
    d->clt_id = conn;
    r->pay_end[r->submitted & mask] = head;
    r->pay_head = head;
    r->submitted++;
    return APUS_OK;
}

static int ring_flush(apus_replica *r)
{
    if (r->flushed == r->submitted) return APUS_OK;
    if (r->cfg.ring_mode == APUS_RING_HOST_MAPPED) {
        __sync_synchronize();                        /* descriptors + payload before the doorbell */
        r->hw->sub_tail = r->submitted;
        r->flushed = r->submitted;
        r->pay_flushed = r->pay_head;
        return APUS_OK;
    }
    DeviceGuard g(r->cfg.device);
    const uint32_t mask = r->ring_slots - 1;
    /* descriptors [flushed, submitted) in at most two runs */
    uint64_t f = r->flushed;
    while (f < r->submitted) {
        uint64_t i0 = f & mask;
        uint64_t run = r->submitted - f;
        if (i0 + run > r->ring_slots) run = r->ring_slots - i0;
        CK(cudaMemcpyAsync(r->ring_desc_dev + i0, r->ring_desc_host + i0, run * sizeof(apus_slot_t),
                           cudaMemcpyHostToDevice, r->copy_stream));
        f += run;
    }
    /* payload bytes [pay_flushed, pay_head) of the byte ring, in at most two runs */
    const uint64_t R = r->ring_bytes;
    uint64_t pf = r->pay_flushed;
    while (pf < r->pay_head) {
        uint64_t p0 = pf % R;
        uint64_t run = r->pay_head - pf;
        if (p0 + run > R) run = R - p0;
        CK(cudaMemcpyAsync(r->ring_pay_dev + p0, r->ring_pay_host + p0, run, cudaMemcpyHostToDevice, r->copy_stream));
        pf += run;
    }
    *r->sub_tail_stage = r->submitted;
    CK(cudaMemcpyAsync(r->sub_tail_dev, r->sub_tail_stage, 8, cudaMemcpyHostToDevice, r->copy_stream));
    CK(cudaStreamSynchronize(r->copy_stream));
    r->flushed = r->submitted;
    r->pay_flushed = r->pay_head;
    return APUS_OK;
}

extern "C" int apus_submit(apus_replica_t *r, uint8_t type, uint16_t connection_id, uint64_t req_id,
                           const void *cmd, uint16_t len, uint64_t *ticket)
{
    if (!r) return fail("null argument");
    if (!is_leader(r)) return fail("submit on a follower (proxy.c:235 only submits when is_leader())");
    if (len && !cmd) return fail("null payload");
    int rc = ring_put(r, type, connection_id, req_id, cmd, len);
    if (rc != APUS_OK) { if (rc == APUS_RETRY) snprintf(g_err, sizeof g_err, "submission ring full"); return rc; }
    if (ticket) *ticket = r->submitted;
    if (!r->defer) return ring_flush(r);
    return APUS_OK;
}

extern "C" int apus_submit_batch(apus_replica_t *r, uint32_t n, const uint8_t *types, const uint16_t *conns,
                                 const uint64_t *req_ids, const uint16_t *lens, const void *payloads,
                                 size_t stride, uint64_t *first_ticket)
{
    if (!r || !types || !conns || !req_ids || !lens) return fail("null argument");
    if (!is_leader(r)) return fail("submit on a follower");
    uint64_t t0 = r->submitted + 1;
    /* all or nothing: remember the ring state */
    const uint64_t save_sub = r->submitted, save_head = r->pay_head;
    for (uint32_t k = 0; k < n; k++) {
        const uint8_t *p = payloads ? (const uint8_t *)payloads + (size_t)k * stride : NULL;
        int rc = ring_put(r, types[k], conns[k], req_ids[k], p, lens[k]);
        if (rc != APUS_OK) {
            r->submitted = save_sub; r->pay_head = save_head;
            if (rc == APUS_RETRY) snprintf(g_err, sizeof g_err, "submission ring full");
            return rc;
        }
    }
    if (first_ticket) *first_ticket = t0;
    if (!r->defer) return ring_flush(r);
    return APUS_OK;
}

extern "C" int apus_submit_defer(apus_replica_t *r, int defer)
{
    if (!r) return fail("null argument");
    r->defer = defer;
    return APUS_OK;
}
extern "C" int apus_submit_flush(apus_replica_t *r)
{
    if (!r) return fail("null argument");
    return ring_flush(r);
}

extern "C" uint64_t apus_committed_tickets(apus_replica_t *r) { return r ? r->hw->committed_tickets : 0; }

extern "C" int apus_progress(apus_replica_t *r, uint64_t *offset, uint64_t *count)
{
    if (!r) return fail("null argument");
    /* count first: the kernel writes the offset before the count */
    const uint64_t c = r->hw->committed_tickets;
    __sync_synchronize();
    if (count) *count = c;
    if (offset) *offset = r->hw->commit_off;
    return r->hw->error ? fail("kernel reported protocol error %llu", (unsigned long long)r->hw->error) : APUS_OK;
}

extern "C" int apus_wait_committed(apus_replica_t *r, uint64_t ticket, int64_t timeout_us)
{
    if (!r) return fail("null argument");
    if (r->hw->committed_tickets >= ticket) return APUS_OK;
    struct timespec t0, t;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    uint32_t spins = 0;
    while (r->hw->committed_tickets < ticket) {
        if ((++spins & 0x3ff) == 0) {
            if (r->hw->error) return fail("kernel reported protocol error %llu", (unsigned long long)r->hw->error);
            if (timeout_us >= 0) {
                clock_gettime(CLOCK_MONOTONIC, &t);
                int64_t us = (t.tv_sec - t0.tv_sec) * 1000000 + (t.tv_nsec - t0.tv_nsec) / 1000;
                if (us > timeout_us) { snprintf(g_err, sizeof g_err, "timeout"); return APUS_RETRY; }
            }
        }
    }
    return APUS_OK;
}

/* Closed loop with ONE request in flight, timed on the host around the two ABI steps a
 * proxy thread performs (enqueue, then spin until committed -- proxy.c:108-161). */
extern "C" int apus_closed_loop(apus_replica_t *r, uint32_t n, uint16_t payload_len, uint16_t connection_id,
                                uint64_t first_req_id, uint32_t *lat_ns)
{
    if (!r || !lat_ns) return fail("null argument");
    if (!is_leader(r)) return fail("submit on a follower");
    uint8_t *buf = (uint8_t *)malloc(payload_len ? payload_len : 1);
    if (!buf) return fail("out of memory");
    for (uint32_t k = 0; k < payload_len; k++) buf[k] = (uint8_t)(k * 131u + 7u);
    struct timespec t0, t1;
    for (uint32_t i = 0; i < n; i++) {
        clock_gettime(CLOCK_MONOTONIC, &t0);
        int rc = ring_put(r, APUS_SEND, connection_id, first_req_id + i, buf, payload_len);
        if (rc == APUS_OK) rc = ring_flush(r);
        if (rc != APUS_OK) { free(buf); return rc == APUS_RETRY ? rc : fail("closed loop: submit failed"); }
        const uint64_t ticket = r->submitted;
        uint32_t spins = 0;
        while (r->hw->committed_tickets < ticket) {
            if ((++spins & 0xffffff) == 0) {
                clock_gettime(CLOCK_MONOTONIC, &t1);
                if (t1.tv_sec - t0.tv_sec > 10 || r->hw->error) { free(buf); return fail("closed loop: commit timeout"); }
            }
        }
        clock_gettime(CLOCK_MONOTONIC, &t1);
        int64_t ns = (t1.tv_sec - t0.tv_sec) * 1000000000ll + (t1.tv_nsec - t0.tv_nsec);
        lat_ns[i] = ns > 0xffffffffll ? 0xffffffffu : (uint32_t)ns;
    }
    free(buf);
    return APUS_OK;
}

/* ---- inspection ---------------------------------------------------------------- */
extern "C" int apus_log_offsets(apus_replica_t *r, apus_log_offsets_t *out)
{
    if (!r || !out) return fail("null argument");
    DeviceGuard g(r->cfg.device);
    apus_loghdr_t h;
    CK(cudaMemcpyAsync(&h, r->region + APUS_HDR_OFF, sizeof h, cudaMemcpyDeviceToHost, r->copy_stream));
    CK(cudaStreamSynchronize(r->copy_stream));
    out->head = h.head; out->apply = h.apply; out->commit = h.commit; out->end = h.end;
    out->tail = h.tail; out->old_end = h.old_end; out->old_commit = h.old_commit; out->len = h.len;
    return APUS_OK;
}

extern "C" int apus_log_read(apus_replica_t *r, uint64_t off, uint64_t len, void *dst)
{
    if (!r || !dst) return fail("null argument");
    if (off + len > r->log_len) return fail("range beyond the log");
    DeviceGuard g(r->cfg.device);
    CK(cudaMemcpyAsync(dst, r->region + r->entries_off + off, len, cudaMemcpyDeviceToHost, r->copy_stream));
    CK(cudaStreamSynchronize(r->copy_stream));
    return APUS_OK;
}

extern "C" int apus_get_stats(apus_replica_t *r, apus_stats_t *out)
{
    if (!r || !out) return fail("null argument");
    DeviceGuard g(r->cfg.device);
    apus_ctrl_t c;
    CK(cudaMemcpyAsync(&c, r->region, sizeof c, cudaMemcpyDeviceToHost, r->copy_stream));
    CK(cudaStreamSynchronize(r->copy_stream));
    memset(out, 0, sizeof *out);
    out->tickets_submitted = r->submitted;
    out->tickets_consumed = c.consumed;
    out->tickets_committed = c.committed_tickets;
    out->entries_acked = c.acked;
    out->bytes_replicated = c.bytes_replicated;
    out->batches = c.batches;
    out->kernel_launches = r->launches;
    out->lat_samples = c.lat_count;
    out->auto_heads = c.auto_heads;
    out->entries_published = c.published;
    for (int i = 0; i < 8; i++) { out->phase_ns[i] = c.phase_ns[i]; out->turn_ns[i] = c.turn_ns[i]; }
    return APUS_OK;
}

extern "C" int apus_latency_samples(apus_replica_t *r, uint32_t *dst, uint32_t max, uint32_t *n)
{
    if (!r || !dst || !n) return fail("null argument");
    DeviceGuard g(r->cfg.device);
    apus_ctrl_t c;
    CK(cudaMemcpyAsync(&c, r->region, sizeof c, cudaMemcpyDeviceToHost, r->copy_stream));
    CK(cudaStreamSynchronize(r->copy_stream));
    uint64_t have = c.lat_count < APUS_LAT_RING ? c.lat_count : APUS_LAT_RING;
    uint32_t take = (uint32_t)(have < max ? have : max);
    uint32_t *tmp = (uint32_t *)malloc(sizeof(uint32_t) * APUS_LAT_RING);
    if (!tmp) return fail("out of memory");
    cudaError_t e = cudaMemcpyAsync(tmp, r->d_lat, sizeof(uint32_t) * APUS_LAT_RING, cudaMemcpyDeviceToHost, r->copy_stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(r->copy_stream);
    if (e != cudaSuccess) { free(tmp); return fail("latency copy failed: %s", cudaGetErrorString(e)); }
    for (uint32_t k = 0; k < take; k++)
        dst[k] = tmp[(c.lat_count - take + k) & (APUS_LAT_RING - 1)];
    free(tmp);
    *n = take;
    return APUS_OK;
}

extern "C" int apus_set_head(apus_replica_t *r, uint64_t head)
{
    if (!r) return fail("null argument");
    if (head >= r->log_len) return fail("head beyond the log");
    DeviceGuard g(r->cfg.device);
    static __thread uint64_t stage;
    stage = head;
    CK(cudaMemcpyAsync(r->region + APUS_HDR_OFF + offsetof(apus_loghdr_t, head), &stage, 8,
                       cudaMemcpyHostToDevice, r->copy_stream));
    CK(cudaStreamSynchronize(r->copy_stream));
    return APUS_OK;
}

extern "C" int apus_remote_apply_offsets(apus_replica_t *r, uint64_t out[APUS_MAX_SERVER_COUNT])
{
    if (!r || !out) return fail("null argument");
    DeviceGuard g(r->cfg.device);
    apus_ctrl_t c;
    CK(cudaMemcpyAsync(&c, r->region, sizeof c, cudaMemcpyDeviceToHost, r->copy_stream));
    CK(cudaStreamSynchronize(r->copy_stream));
    for (int i = 0; i < APUS_MAX_SERVER_COUNT; i++) out[i] = c.apply_off[i];
    return APUS_OK;
}
