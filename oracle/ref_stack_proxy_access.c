/*
 * oracle/ref_stack_proxy_access.c -- TEST INFRASTRUCTURE ONLY (part of oracle/_ref/libref_stack.so).
 * What the reference's proxy.c and db-interface.c counted while the stack ran: requests released to the application
 * (update_state calls, proxy.c:263-267), requests admitted (proxy.c:115), and the bytes handed to BerkeleyDB by
 * store_cmd (db-interface.c:81).  Separate from ref_stack_access.c because proxy.h and dare's debug.h do not mix.
 */
#include <stdint.h>
#include <proxy/proxy.h>

extern uint32_t records_len;
uint64_t refstack_highest_rec(void *proxy) { return ((proxy_node *)proxy)->highest_rec; }
uint64_t refstack_cur_rec(void *proxy) { return ((proxy_node *)proxy)->cur_rec; }
uint32_t refstack_records_len(void) { return records_len; }
