#!/bin/bash
# One GPU: dint_udp_server (the reference's UDP server shape over the C ABI) behind the same multi-socket loopback
# replayer that times the unmodified reference server (cpu_baseline.udp_as_shipped).  Prints the replayer's JSON.
set -u
cd "$(dirname "$0")/.."
PORT=${PORT:-20999}
mkdir -p gpurun_out
python - <<'PY'
import sys
sys.path[:0] = ["tests", "."]
import trace_gen as T
T.fasst_random(1 << 20, 24000000, seed=1, weights=(0.6, 0.15, 0.05, 0.2)).tofile("gpurun_out/udp_trace.bin")
PY
dint_b200/lib/dint_udp_server lock_fasst --bind 127.0.0.1 --port "$PORT" &
SRV=$!
sleep 12                                   # CUDA context + tables
for cfg in "8 32" "32 64" "64 256"; do
  set -- $cfg
  oracle/_ref/udp_blast gpurun_out/udp_trace.bin 9 "$PORT" "$1" "$2" 4
done
kill "$SRV"                                # exactly the server we started
wait "$SRV" 2>/dev/null
rm -f gpurun_out/udp_trace.bin
