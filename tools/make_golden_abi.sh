#!/bin/sh
# Regenerates tests/golden/abi_reference.txt from the REFERENCE headers (needs /root/reference).
set -e
cd "$(dirname "$0")/.."
make -s -C oracle _ref/dropin_host_ref
oracle/_ref/dropin_host_ref abi > tests/golden/abi_reference.txt
wc -l tests/golden/abi_reference.txt
