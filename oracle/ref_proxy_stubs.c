/*
 * oracle/ref_proxy_stubs.c -- TEST DOUBLES linked next to the reference's UNMODIFIED
 * src/proxy/proxy.c in oracle/_ref/libref_proxy.so (TEST INFRASTRUCTURE ONLY).
 *
 * proxy.c's two side libraries need BerkeleyDB and libconfig headers that this image does
 * not have, and neither is on the replication hot path.  They are replaced by
 * recording stand-ins so that the drop-in test can check WHAT proxy.c asks them to do:
 *   src/config-comp/config-proxy.c:14-45  proxy_read_config  -> values from the environment
 *   src/db/db-interface.c:22-128          initialize_db / store_record / dump_records /
 *                                         get_records_len / close_db -> in-memory record log
 */
#include <arpa/inet.h>
#include <netinet/in.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <proxy/proxy.h>
#include <config-comp/config-proxy.h>

struct db_t { uint32_t n; uint64_t bytes; uint32_t sizes[1 << 20]; uint8_t *blob; uint64_t cap; uint32_t dumps, loads; };
static struct db_t g_db;

db *initialize_db(const char *db_name, uint32_t flag) { (void)db_name; (void)flag; memset(&g_db, 0, sizeof g_db); return &g_db; }
void close_db(db *d, uint32_t mode) { (void)d; (void)mode; }
int store_record(db *d, size_t data_size, void *data)
{
    if (d->n < (1u << 20)) d->sizes[d->n] = (uint32_t)data_size;
    if (d->bytes + data_size > d->cap) { d->cap = (d->cap ? d->cap * 2 : (1u << 20)) + data_size; d->blob = (uint8_t *)realloc(d->blob, d->cap); }
    memcpy(d->blob + d->bytes, data, data_size);          /* records in append order, like a BDB recno cursor walk */
    d->n++; d->bytes += data_size;
    return 0;
}
/* db-interface.c:98-128: every record's bytes, in order, back to back */
void dump_records(db *d, void *buf) { memcpy(buf, d->blob, d->bytes); d->dumps++; }
uint32_t stub_db_dumps(void) { return g_db.dumps; }
uint32_t get_records_len() { return (uint32_t)g_db.bytes; }

/* inspection for the test driver */
uint32_t stub_db_count(void) { return g_db.n; }
uint32_t stub_db_size(uint32_t i) { return g_db.sizes[i]; }

int proxy_read_config(struct proxy_node_t *cur_node, const char *config_path)
{
    (void)config_path;
    const char *port = getenv("stub_port");
    cur_node->db_name = "stub_db";
    cur_node->req_log = 0;
    memset(&cur_node->sys_addr, 0, sizeof cur_node->sys_addr);
    cur_node->sys_addr.s_addr.sin_family = AF_INET;
    inet_pton(AF_INET, "127.0.0.1", &cur_node->sys_addr.s_addr.sin_addr);
    cur_node->sys_addr.s_addr.sin_port = htons((uint16_t)(port ? atoi(port) : 8888));
    cur_node->sys_addr.s_sock_len = sizeof(cur_node->sys_addr.s_addr);
    return 0;
}

/* what the test reads back from proxy_node_t */
uint64_t stub_highest_rec(struct proxy_node_t *p) { return p->highest_rec; }
uint64_t stub_cur_rec(struct proxy_node_t *p) { return p->cur_rec; }

/* the application-side driver of oracle/app_driver.inc on THIS proxy (the reference's unmodified proxy.c linked on the
 * GPU engine): the same closed-loop load the reference arm runs on its own stack (refstack_drive) */
#define refstack_ragged_len stub_ragged_len
#include "app_driver.inc"
int stub_drive(void *proxy, int threads, int nconn, int64_t nreq, int plen, uint64_t *lat_ns, double *seconds, int accept_close)
{
    return app_drive(proxy, threads, nconn, nreq, plen, lat_ns, seconds, accept_close);
}
