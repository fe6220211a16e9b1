#!/bin/bash
# benchmarks/run_gpu.sh -- single-box launcher of an APUS group on the GPU engine: the counterpart of the reference's
# benchmarks/run.sh:23-81 (StartDare / FindLeader / StartBenchmark / StopDare) and reconf_bench.sh:249-343 (leader-kill
# drill) for ONE 8xB200 box instead of a cluster reached over ssh.  The operator surface is the reference's:
#   server_type / server_idx / group_size / config_path / dare_log_file   environment of every replica (run.sh:26)
#   LD_PRELOAD=<interpose.so>                                             the unmodified interposer, linked on libapus_dare.so
#   one libconfig file per replica (port / db_name are per file, target/nodes.local.cfg:5,11)
#   "] LEADER" in srv<i>.log                                              how run.sh:52 finds the leader
# plus, per replica, apus_gpu=<i> (one GPU per replica) and a shared apus_rendezvous directory (CUDA IPC handles).
#
#   benchmarks/run_gpu.sh --app=redis --scount=5 --ccount=16 --rcount=200000 [--dsize=128] [--kill-leader]
#   benchmarks/run_gpu.sh --app=memcached --scount=7 --ccount=16 --rcount=100000 --dsize=1024        (BASELINE config 4)
#   APP_CMD="memcached -p %PORT%" benchmarks/run_gpu.sh --app=custom --scount=7     (any server the interposer can wrap)
set -u
HERE="$(cd "$(dirname "$0")/.." && pwd)"
REFBIN="${REFBIN:-$HERE/oracle/_ref}"
INTERPOSE="${INTERPOSE:-$REFBIN/interpose.so}"
APP=redis; server_count=3; client_count=16; request_count=100000; dsize=128; kill_leader=0; base_port=8888
for arg in "$@"; do
  case $arg in
    --app=*) APP="${arg#*=}";; --scount=*) server_count="${arg#*=}";; --ccount=*) client_count="${arg#*=}";;
    --rcount=*) request_count="${arg#*=}";; --dsize=*) dsize="${arg#*=}";; --port=*) base_port="${arg#*=}";;
    --kill-leader) kill_leader=1;;
    *) echo "usage: $0 --app=redis|memcached|custom --scount=N --ccount=C --rcount=R [--dsize=B] [--port=P] [--kill-leader]"; exit 1;;
  esac
done
ngpu=$(nvidia-smi -L 2>/dev/null | wc -l); [ "$ngpu" -lt 1 ] && { echo "no GPU visible: the engine has no CPU fallback"; exit 1; }
RUN="$(mktemp -d /tmp/apus-run-XXXXXX)"
declare -a pids
if [ "$server_count" -gt "$ngpu" ]; then
  # fewer GPUs than replicas: the processes' contexts are time-sliced (milliseconds per turn), a heartbeat timeout sized for
  # resident kernels would fire spuriously -- functional run only, the latencies mean nothing (tools/failover_drill.py does the same)
  export apus_hb_period_us=${apus_hb_period_us:-2000} apus_hb_timeout_us=${apus_hb_timeout_us:-400000} apus_elec_timeout_us=${apus_elec_timeout_us:-100000,300000}
  echo "note: $server_count replicas on $ngpu GPU(s): time-sliced contexts, relaxed failure-detector timeouts"
fi

StartDare() {
  for ((i=0; i<$1; ++i)); do
    mkdir -p "$RUN/node$i"
    cat > "$RUN/node$i/node.cfg" <<CFG
db_name = "node_test$i";
req_log = 0;
ip_address = "127.0.0.1";
port = $((base_port + i));
dare_global_config = { hb_period = ${HB_PERIOD:-0.001}; elec_timeout_low = ${ELEC_LOW:-10000}; elec_timeout_high = ${ELEC_HIGH:-30000};
                       retransmit_period = 0.02; rc_info_period = 0.01; log_pruning_period = 0.03; };
CFG
    if [ "$APP" = redis ]; then run_dare=( "$REFBIN/redis-server" --port $((base_port + i)) --save "" --bind 127.0.0.1 )
    elif [ "$APP" = memcached ]; then run_dare=( "$REFBIN/memcached" -u root -p $((base_port + i)) -U 0 -t 4 -l 127.0.0.1 -m 1024 )
    else run_dare=( ${APP_CMD//%PORT%/$((base_port + i))} ); fi
    pushd "$RUN/node$i" > /dev/null
    env server_type=start server_idx=$i group_size=$1 config_path="$RUN/node$i/node.cfg" \
        dare_log_file="$RUN/srv$i.log" apus_gpu=$((i % ngpu)) apus_rendezvous="$RUN/rdv" LD_PRELOAD="$INTERPOSE" \
        nohup "${run_dare[@]}" > app.out 2>&1 &
    pids[$i]=$!
    popd > /dev/null
  done
  echo -e "\tinitial servers: p0..p$(($1 - 1)) on $ngpu GPU(s), PIDs: ${pids[*]}"
}

StopDare() { for p in "${pids[@]}"; do kill -2 "$p" 2>/dev/null; done; sleep 1; for p in "${pids[@]}"; do kill -9 "$p" 2>/dev/null; done; }

FindLeader() {          # the latest "[T<term>] LEADER" line over all logs (run.sh:46-72)
  leader_idx=""; max_term=0
  for ((i=0; i<server_count; ++i)); do
    [ -f "$RUN/srv$i.log" ] || continue
    while read -r term; do
      [ -n "$term" ] && [ "$term" -gt "$max_term" ] && { max_term=$term; leader_idx=$i; }
    done < <(grep "\] LEADER" "$RUN/srv$i.log" | sed -E 's/.*\[T([0-9]+)\] LEADER.*/\1/')
  done
  [ -n "$leader_idx" ] && echo "Leader: p$leader_idx (term $max_term)"
}

StartBenchmark() {
  if [ "$APP" = memcached ]; then   # config 4: set/get half and half (apps/memcached/run uses memslap; here a Python client)
    python "$HERE/tests/memcached_group.py" $((base_port + leader_idx)) "$client_count" $((request_count / client_count / 2)) "$dsize"
  else
    "$REFBIN/redis-benchmark" -t set -d "$dsize" -p $((base_port + leader_idx)) -n "$request_count" -c "$client_count" -r 100000 -q
  fi
}

trap 'StopDare; echo "logs kept in $RUN"' EXIT
StartDare "$server_count"
for t in $(seq 1 600); do FindLeader > /dev/null; [ -n "$leader_idx" ] && break; sleep 0.1; done
FindLeader || { echo "no leader after 60 s; see $RUN/srv*.log"; exit 1; }
sleep 1
if [ "$kill_leader" = 1 ]; then
  StartBenchmark > "$RUN/clt1.log" 2>&1 &
  sleep 2
  old=$leader_idx; t_kill=$(date +%s.%N)
  kill -9 "${pids[$old]}"; echo "killed the leader p$old (PID ${pids[$old]})"
  for t in $(seq 1 2000); do FindLeader > /dev/null; [ "$leader_idx" != "$old" ] && break; sleep 0.005; done
  t_new=$(date +%s.%N)
  FindLeader; echo "recovery (kill -> next \"] LEADER\" line, polled every 5 ms): $(awk -v a="$t_new" -v b="$t_kill" 'BEGIN { printf "%.1f", (a - b) * 1000 }') ms"
  wait; sleep 1
  StartBenchmark
else
  StartBenchmark
fi
