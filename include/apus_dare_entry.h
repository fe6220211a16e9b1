/*
 * apus_dare_entry.h -- the "engine entry" surface APUS's kept files link against.
 *
 * src/proxy/proxy.c (kept unmodified) needs exactly these things from
 * the consensus engine, all declared in the reference by
 *     src/include/dare/dare_server.h:142-160, 197-203   dare_server_input_t, dare_server_init,
 *                                                       dare_server_shutdown, is_leader, get_node_id
 *     src/include/dare/message.h:5-22                   tailq_cmd_t, tailq_entry_t, tailhead, tailq_lock
 *     src/include/dare/dare_sm.h:40-47                  the six proxy callback types
 * This header restates those declarations (same names, same layouts, same meaning) so that
 * libapus_dare.so -- apus_b200/csrc/dare_entry.c, implemented on top of include/apus_gpu.h --
 * can stand in for the reference's libdare.a.  Nothing below the engine entry (ibverbs queue
 * pairs, UD bootstrap, libev loop) exists any more.
 *
 * Configuration comes from the same channels as in the reference: environment variables
 * server_idx / group_size / server_type / dare_log_file / config_path are read by proxy.c
 * (proxy.c:22-89) and arrive here through dare_server_input_t.  Additional, optional variables
 * (all have defaults, old launch scripts keep working):
 *     apus_gpu         CUDA device ordinal of this replica   (default: server_idx % device count)
 *     apus_leader      index of the leader replica           (default: 0; static until the
 *                                                             election control plane lands)
 *     apus_rendezvous  directory used to exchange the 128-byte peer handles between the
 *                      replica processes                     (default: /tmp/apus-rdv-<uid>)
 *     apus_log_size    bytes of entries[]                    (default: LOG_SIZE, 64 MiB)
 */
#ifndef APUS_DARE_ENTRY_H
#define APUS_DARE_ENTRY_H

#include <pthread.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <sys/queue.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- message.h:5-22 -------------------------------------------------------------- */
#ifndef MESSAGE_H
struct tailq_cmd_t {
    uint16_t len;
    uint8_t  cmd[87380];
};
typedef struct tailq_cmd_t tailq_cmd_t;

struct tailq_entry_t {
    uint8_t  type;            /* CONNECT 4 / SEND 5 / CLOSE 6 (proxy.h:9-11) */
    uint16_t connection_id;
    uint64_t req_id;
    tailq_cmd_t cmd;
    TAILQ_ENTRY(tailq_entry_t) entries;
};
typedef struct tailq_entry_t tailq_entry_t;

/* the reference DEFINES these two in the header (tentative definitions, -fcommon);
 * libapus_dare.so carries the one real definition */
TAILQ_HEAD(apus_tailhead_t, tailq_entry_t);
extern struct apus_tailhead_t tailhead;
extern pthread_spinlock_t tailq_lock;
#endif

/* ---- dare_sm.h:40-47 ---------------------------------------------------------------- */
#ifndef DARE_SM_H
typedef void (*proxy_store_cmd_cb_t)(void *data, void *arg);
typedef void (*proxy_do_action_cb_t)(uint16_t clt_id, uint8_t type, size_t data_size, void *data, void *arg);
typedef void (*proxy_create_db_snapshot_cb_t)(void *snapshot, void *arg);
typedef uint32_t (*proxy_get_db_size_cb_t)(void *arg);
typedef int (*proxy_apply_db_snapshot_cb_t)(void *snapshot, uint32_t size, void *arg);
typedef void (*proxy_update_state_cb_t)(void *arg);
#endif

/* ---- dare_server.h:142-160 ------------------------------------------------------------ */
#ifndef DARE_SERVER_H
#define SRV_TYPE_START 1
#define SRV_TYPE_JOIN  2
struct dare_server_input_t {
    FILE   *log;
    char   *name;
    char   *output;
    uint8_t srv_type;
    uint8_t sm_type;
    uint8_t group_size;
    uint8_t server_idx;
    proxy_do_action_cb_t          do_action;
    proxy_store_cmd_cb_t          store_cmd;
    proxy_create_db_snapshot_cb_t create_db_snapshot;
    proxy_get_db_size_cb_t        get_db_size;
    proxy_apply_db_snapshot_cb_t  apply_db_snapshot;
    proxy_update_state_cb_t       update_state;
    char    config_path[128];
    void   *up_para;
};
typedef struct dare_server_input_t dare_server_input_t;

/* dare_server.h:197-203.  dare_server_init is a pthread start routine: it takes ownership of
 * (and frees) its dare_server_input_t, brings the replica up, and then pumps:
 *   leader   drains tailhead FIFO under tailq_lock exactly as get_tailq_message does
 *            (dare_ibv_ud.c:780-790), frees each node, calls store_cmd once per appended
 *            entry and update_state once per committed CSM-type entry, in log order;
 *   follower calls store_cmd and do_action once per committed entry, in log order.
 * All callbacks run on this thread only. */
void   *dare_server_init(void *arg);
void    dare_server_shutdown(void);
int     is_leader(void);
uint8_t get_node_id(void);
#endif

#ifdef __cplusplus
}
#endif
#endif /* APUS_DARE_ENTRY_H */
