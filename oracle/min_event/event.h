/*
 * oracle/min_event/event.h -- TEST INFRASTRUCTURE ONLY (BASELINE config 4).
 * The image carries the libevent 2.1 RUNTIME library (libevent_core-2.1.so.7) but no headers; memcached 1.4.21
 * (apps/memcached/memcached-1.4.21.tar.gz, built unmodified by oracle/build_refapp.sh) needs eight calls of the libevent
 * 1.4-style API and embeds `struct event` in its own structures, reading `ev_base` directly (memcached.c:3889).  This
 * header declares exactly that subset with the 2.1 layout; oracle/min_event/probe.c checks size and field offsets against
 * the library at build time (event_get_struct_event_size, event_base_set) and the build refuses to go on if they differ.
 */
#ifndef APUS_MIN_EVENT_H
#define APUS_MIN_EVENT_H
#include <sys/time.h>
#define EV_TIMEOUT 0x01
#define EV_READ    0x02
#define EV_WRITE   0x04
#define EV_SIGNAL  0x08
#define EV_PERSIST 0x10
#define EVLOOP_ONCE 0x01
#define EVLOOP_NONBLOCK 0x02
struct event_base;
struct event {                      /* libevent 2.1, LP64: 128 bytes; ev_fd at 56, ev_base at 64 (event2/event_struct.h) */
    char ev_evcallback[40];
    char ev_timeout_pos[16];
    int ev_fd;
    struct event_base *ev_base;
    char ev_rest[56];
};
struct event_base *event_init(void);
void event_set(struct event *, int, short, void (*)(int, short, void *), void *);
int event_base_set(struct event_base *, struct event *);
int event_add(struct event *, const struct timeval *);
int event_del(struct event *);
int event_base_loop(struct event_base *, int);
const char *event_get_version(void);
size_t event_get_struct_event_size(void);
#define evtimer_set(ev, cb, arg) event_set((ev), -1, 0, (cb), (arg))
#define evtimer_add(ev, tv) event_add((ev), (tv))
#define evtimer_del(ev) event_del(ev)
#endif
