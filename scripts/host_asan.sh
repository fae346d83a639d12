#!/bin/bash
# ASan/UBSan build of the host side (shim + text VCF reader + the test-only device stand-in) linked with the reference's
# objects as built by pbwt_b200/host/Makefile; runs the CPU host test-suite against it.  Needs /root/reference at build time.
set -euo pipefail
ROOT=$(cd "$(dirname "$0")/.." && pwd)
REF=${PBWT_REF:-/root/reference}
OUT=$(mktemp -d)
HOST=$ROOT/pbwt_b200/host
SAN="-g -O1 -fsanitize=address,undefined -w"
make -C "$HOST" > /dev/null
make -C "$ROOT/oracle" port > /dev/null
gcc $SAN -I"$REF" -I"$HOST/_build" -I"$ROOT/include" -c "$HOST/pbwt_gpu_shim.c" -o "$OUT/shim.o"
gcc $SAN -I"$REF" -I"$HOST/_build" -I"$ROOT/include" -c "$HOST/vcf_text.c" -o "$OUT/vcf.o"
gcc $SAN -I"$ROOT/include" -I"$ROOT/oracle" -c "$ROOT/tests/host_mock/mock_pbwtgpu.c" -o "$OUT/mock.o"
gcc -fsanitize=address,undefined -o "$OUT/hostcheck_asan" \
    $(ls "$HOST"/_build/*.o | grep -v "shim_selftest.o\|pbwt_gpu_shim.o\|vcf_text.o") "$OUT/shim.o" "$OUT/vcf.o" "$OUT/mock.o" \
    -L"$ROOT/oracle" -lpbwt_oracle -Wl,-rpath,"$ROOT/oracle" -lz -lm
cd "$ROOT"
ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 UBSAN_OPTIONS=halt_on_error=1 PBWT_HOSTCHECK="$OUT/hostcheck_asan" \
    python -m pytest tests/test_host_shim_cpu.py -q -x
