/* oracle/ref_stubs/ev.h -- test double: the one declaration of libev that
 * src/include/dare/dare_server.h needs in order to parse (struct ev_loop *loop).
 * Used only to compile the reference's proxy.c for tests/test_gpu_proxy_dropin.py. */
#ifndef APUS_STUB_EV_H
#define APUS_STUB_EV_H
struct ev_loop;
#endif
