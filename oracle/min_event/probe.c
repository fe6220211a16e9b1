/* oracle/min_event/probe.c -- build-time check of event.h against the libevent runtime library (see event.h) */
#include <stddef.h>
#include <stdio.h>
#include <event.h>
static void cb(int a, short b, void *c) { (void)a; (void)b; (void)c; }
int main(void)
{
    struct event_base *b = event_init();
    struct event ev;
    event_set(&ev, 5, EV_READ | EV_PERSIST, cb, NULL);
    event_base_set(b, &ev);
    printf("%s size %zu (mine %zu) fd %d base ok %d\n", event_get_version(), event_get_struct_event_size(), sizeof ev, ev.ev_fd, ev.ev_base == b);
    return !(event_get_struct_event_size() == sizeof ev && ev.ev_fd == 5 && ev.ev_base == b);
}
