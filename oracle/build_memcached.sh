#!/bin/bash
# oracle/build_memcached.sh -- TEST INFRASTRUCTURE ONLY (BASELINE config 4).
# Builds oracle/_ref/memcached from the tarball where it lies under $REF (apps/memcached/memcached-1.4.21.tar.gz, the
# version apps/memcached/mk builds), unmodified.  memcached needs libevent: the image has the 2.1 runtime library but no
# headers, so the build goes against oracle/min_event/event.h (the eight calls memcached uses) after probe.c has checked
# that header against the library.  Flags added to the stock build: -fcommon (gcc >= 10: `hash` is defined in a header),
# no -Werror.  Nothing is copied into the repository; scratch goes to oracle/_ref/build_mc and is removed.
set -e
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
TGZ=$REF/apps/memcached/memcached-1.4.21.tar.gz
if [ ! -f "$TGZ" ]; then echo "reference tree absent: keeping prebuilt oracle/_ref/memcached (if any)"; exit 0; fi
if [ -x "$OUT/memcached" ] && [ "$OUT/memcached" -nt "$HERE/min_event/event.h" ] && [ -z "$FORCE" ]; then echo "oracle/_ref/memcached up to date"; exit 0; fi
LIBEV=$(ls /usr/lib/x86_64-linux-gnu/libevent_core-2.1.so.* /lib/x86_64-linux-gnu/libevent_core-2.1.so.* 2>/dev/null | head -1)
if [ -z "$LIBEV" ]; then echo "no libevent runtime library on this box: memcached not built"; exit 0; fi
mkdir -p "$OUT"
B=$OUT/build_mc
rm -rf "$B"; mkdir -p "$B/prefix/include" "$B/prefix/lib"; cd "$B"
cp "$HERE/min_event/event.h" prefix/include/
ln -s "$LIBEV" prefix/lib/libevent.so
gcc -Iprefix/include "$HERE/min_event/probe.c" -o probe -Lprefix/lib -levent -Wl,-rpath,"$(dirname "$LIBEV")"
./probe || { echo "oracle/min_event/event.h does not match $LIBEV: memcached not built"; exit 1; }
tar xzf "$TGZ"
cd memcached-1.4.21
./configure --with-libevent="$B/prefix" --disable-docs --disable-coverage > ../configure.log 2>&1
make -j${J:-8} memcached CFLAGS="-g -O2 -pthread -w -fcommon" > ../make.log 2>&1 || { tail -20 ../make.log; exit 1; }
cp memcached "$OUT/memcached"
cd "$OUT"; rm -rf "$B"
echo "built oracle/_ref/memcached (1.4.21, $(basename "$LIBEV"))"
