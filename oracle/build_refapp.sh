#!/bin/bash
# oracle/build_refapp.sh -- TEST INFRASTRUCTURE ONLY (drop-in test of SURVEY.md s8f row N2).
#
# Builds, from the tarballs and sources where they lie under $REF (nothing is copied into the repository):
#   oracle/_ref/redis-server, redis-benchmark, redis-cli   apps/redis/redis-2.8.17.tar.gz   (apps/redis/mk)
#   libconfig 1.4.9, BerkeleyDB 5.1.29 (static)             utils/dep-lib/*.tar.gz           (utils/mk)
#   libev 4.15                                             utils/dep-lib/libev-4.15.tar.gz  (utils/mk)
#   oracle/_ref/interpose.so = the reference's UNMODIFIED src/spec_hooks.cpp, src/proxy/proxy.c,
#       src/db/db-interface.c, src/config-comp/config-proxy.c, linked per INTEGRATION.md section 2:
#       libapus_dare.so + libapus_gpu.so in place of libdare.a -lev -libverbs (target/makefile:19).
#   oracle/_ref/libref_stack.so = the reference's COMPLETE software stack, unmodified: src/dare/*.c (election,
#       heartbeats, replication, commit, pruning), utils/rbtree, proxy.c, db-interface.c, config-*.c, on
#       oracle/verbs_shim (a stand-in NIC: process_vm_writev + Unix datagrams) -- "reference-on-shim", SURVEY.md s8d.
#       It pins the oracle's cluster restatement and is the CPU reference arm of bench.py.
# Scratch goes to oracle/_ref/build and is removed at the end; oracle/_ref is git-ignored and travels to the GPU box.
set -e
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
ENGINE=$(cd "$HERE/../apus_b200" && pwd)
J=${J:-8}
if [ ! -f "$REF/apps/redis/redis-2.8.17.tar.gz" ]; then
  echo "reference tree absent: keeping prebuilt oracle/_ref application binaries (if any)"; exit 0
fi
if [ -x "$OUT/redis-server" ] && [ -f "$OUT/interpose.so" ] && [ "$OUT/interpose.so" -nt "$REF/src/proxy/proxy.c" ] \
   && [ -f "$OUT/libref_stack.so" ] && [ "$OUT/libref_stack.so" -nt "$HERE/verbs_shim/verbs_shim.c" ] \
   && [ "$OUT/libref_stack.so" -nt "$HERE/ref_stack_access.c" ] && [ "$OUT/libref_stack.so" -nt "$HERE/ref_stack_proxy_access.c" ] && [ -f "$OUT/interpose_ref.so" ] && [ -f "$OUT/libref_stack_O2.so" ] && [ -z "$FORCE" ]; then
  echo "oracle/_ref application binaries up to date"; exit 0
fi
B=$OUT/build
rm -rf "$B"; mkdir -p "$B"; cd "$B"
tar xzf "$REF/apps/redis/redis-2.8.17.tar.gz"
make -C redis-2.8.17 -j$J MALLOC=libc > redis.log 2>&1
cp redis-2.8.17/src/redis-server redis-2.8.17/src/redis-benchmark redis-2.8.17/src/redis-cli "$OUT/"
tar xzf "$REF/utils/dep-lib/libconfig-1.4.9.tar.gz"
(cd libconfig-1.4.9 && ./configure --disable-shared --disable-cxx --with-pic > ../libconfig.log 2>&1 \
   && make -j$J >> ../libconfig.log 2>&1)
tar xzf "$REF/utils/dep-lib/db-5.1.29.tar.gz"
(cd db-5.1.29/build_unix && ../dist/configure --disable-shared --with-pic --disable-cxx --disable-java --disable-tcl \
   --disable-replication > ../../bdb.log 2>&1 && make -j$J libdb.a >> ../../bdb.log 2>&1)
tar xzf "$REF/utils/dep-lib/libev-4.15.tar.gz"
(cd libev-4.15 && ./configure --disable-shared --with-pic > ../libev.log 2>&1 && make -j$J >> ../libev.log 2>&1)
INC="-I$HERE/ref_stubs -I$REF/src/include -I$REF/src -I$B/libconfig-1.4.9/lib -I$B/db-5.1.29/build_unix"
CF="-fPIC -rdynamic -O0 -g -w -fcommon -DDEBUG=0"          # the flags of target/src/*/subdir.mk (+ -fcommon)
gcc $CF -std=gnu99 $INC -c "$REF/src/proxy/proxy.c" -o proxy.o
gcc $CF -std=gnu99 $INC -c "$REF/src/db/db-interface.c" -o db-interface.o
gcc $CF -std=gnu99 $INC -c "$REF/src/config-comp/config-proxy.c" -o config-proxy.o
g++ -fPIC -rdynamic -O0 -g -w -I"$REF/src" -c "$REF/src/spec_hooks.cpp" -o spec_hooks.o
g++ -shared -Wl,-soname,interpose.so -o "$OUT/interpose.so" spec_hooks.o proxy.o db-interface.o config-proxy.o \
    libconfig-1.4.9/lib/.libs/libconfig.a db-5.1.29/build_unix/libdb.a \
    -L"$ENGINE" -lapus_dare -lapus_gpu -Wl,-rpath,'$ORIGIN/../../apus_b200' -lpthread -ldl -lm
# the complete reference stack on the verbs shim; flags of target/src/dare/subdir.mk (+ -fcommon, -g): -O0 as the
# reference builds, and a second time with -O2 ("build twice") -> libref_stack_O2.so
DINC="-I$REF/src/include/dare -I$REF/utils/rbtree/include -I$HERE/verbs_shim -I$B/libev-4.15"
SINC="-I$B/libev-4.15 -I$HERE/verbs_shim -I$REF/src/include -I$REF/src -I$B/libconfig-1.4.9/lib -I$B/db-5.1.29/build_unix"
for OPT in O0 O2; do
  S=stack; [ $OPT = O2 ] && S=stack_O2
  mkdir -p $S
  for f in "$REF"/src/dare/*.c "$REF"/utils/rbtree/src/*.c; do
    gcc -fPIC -rdynamic -std=gnu99 -$OPT -g -w -fcommon $DINC -c "$f" -o "$S/$(basename "$f" .c).o"
  done
  # proxy.c stays at -O0 in both: its commit wait `while (cur_rec > proxy->highest_rec);` (proxy.c:160) reads a plain
  # field another thread updates; at -O2 the load is hoisted and the application thread never wakes up
  gcc -fPIC -rdynamic -O0 -g -w -fcommon -DDEBUG=0 -std=gnu99 $SINC -c "$REF/src/proxy/proxy.c" -o $S/proxy.o
  gcc -fPIC -rdynamic -$OPT -g -w -fcommon -DDEBUG=0 -std=gnu99 $SINC -c "$REF/src/config-comp/config-dare.c" -o $S/config-dare.o
  gcc -fPIC -rdynamic -$OPT -g -w -fcommon -DDEBUG=0 -std=gnu99 $INC -c "$REF/src/db/db-interface.c" -o $S/db-interface.o
  gcc -fPIC -rdynamic -$OPT -g -w -fcommon -DDEBUG=0 -std=gnu99 $INC -c "$REF/src/config-comp/config-proxy.c" -o $S/config-proxy.o
  gcc -fPIC -O2 -g -std=gnu99 -Wall -I"$HERE/verbs_shim" -c "$HERE/verbs_shim/verbs_shim.c" -o $S/verbs_shim.o
  gcc -fPIC -$OPT -g -std=gnu99 -w -fcommon $DINC -c "$HERE/ref_stack_access.c" -o $S/ref_stack_access.o
  gcc -fPIC -$OPT -g -std=gnu99 -w -fcommon $SINC -c "$HERE/ref_stack_proxy_access.c" -o $S/ref_stack_proxy_access.o
done
gcc -shared -o "$OUT/libref_stack.so" stack/*.o \
    libev-4.15/.libs/libev.a libconfig-1.4.9/lib/.libs/libconfig.a db-5.1.29/build_unix/libdb.a -lpthread -lm
gcc -shared -o "$OUT/libref_stack_O2.so" stack_O2/*.o \
    libev-4.15/.libs/libev.a libconfig-1.4.9/lib/.libs/libconfig.a db-5.1.29/build_unix/libdb.a -lpthread -lm
# the reference's interposer on the reference's OWN stack (shim NIC): the CPU-side counterpart of interpose.so, used to
# run the same redis drop-in scenario against the reference itself (tests/test_refstack_redis.py)
g++ -shared -Wl,-soname,interpose.so -o "$OUT/interpose_ref.so" spec_hooks.o stack/*.o \
    libev-4.15/.libs/libev.a libconfig-1.4.9/lib/.libs/libconfig.a db-5.1.29/build_unix/libdb.a -lpthread -ldl -lm
cd "$OUT"; rm -rf "$B"
echo "built oracle/_ref/{redis-server,redis-benchmark,redis-cli,interpose.so,interpose_ref.so,libref_stack.so,libref_stack_O2.so}"
