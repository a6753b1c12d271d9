#!/bin/bash
# Regenerates tests/golden/dropin_reference.jsonl and dropin_domains.jsonl by compiling
# tests/cpp/dropin_scenarios.cpp against the UNMODIFIED reference header where it lies
# (/root/reference/dequan.h; nothing is copied).  Run in the build container only.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF_DIR="${REF_DIR:-/root/reference}"
TMP="$(mktemp -d)"
g++ -std=c++11 -O2 -DNDEBUG -I"$REF_DIR" -o "$TMP/ref_scen" "$HERE/../cpp/dropin_scenarios.cpp"
"$TMP/ref_scen" > "$HERE/dropin_reference.jsonl"
"$TMP/ref_scen" domains > "$HERE/dropin_domains.jsonl"
rm -rf "$TMP"
wc -l "$HERE/dropin_reference.jsonl" "$HERE/dropin_domains.jsonl"
