#!/bin/bash
# latency-plan sweep (not a test): GPAD_LATENCY_PLAN / THREADS variants per problem size
run() { echo "--- plan=${GPAD_LATENCY_PLAN:-auto} threads=${GPAD_LATENCY_THREADS:-auto} $*"; python tests/prof_latency.py "$@" 30 | tail -1; }
for plan in "" "cluster:8" "cluster:4" "grid:16" "grid:32"; do GPAD_LATENCY_PLAN=$plan run 10 15; done
for plan in "" "grid:74" "grid:37" "grid:120"; do for th in "" 256; do GPAD_LATENCY_PLAN=$plan GPAD_LATENCY_THREADS=$th run 30 30; done; done
for plan in "" "grid:74" "grid:37" "grid:120"; do for th in "" 256; do GPAD_LATENCY_PLAN=$plan GPAD_LATENCY_THREADS=$th run 10 100; done; done
