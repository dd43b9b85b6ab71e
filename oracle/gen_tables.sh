#!/bin/sh
# Regenerates brotli_b200/data/brotli_tables.bin from the reference's table objects.
set -e
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
gcc -O1 -I$REF/c -I$REF/c/include -o /tmp/gen_tables $HERE/gen_tables.c \
    $REF/c/common/dictionary.c $REF/c/common/context.c $REF/c/common/platform.c \
    $REF/c/enc/dictionary_hash.c
/tmp/gen_tables $HERE/../brotli_b200/data/brotli_tables.bin
ls -l $HERE/../brotli_b200/data/brotli_tables.bin
