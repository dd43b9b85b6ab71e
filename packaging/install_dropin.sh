#!/bin/bash
# Lays the library out the way consumers of the reference find libbrotlienc:
#   <prefix>/lib/libbrotlienc.so -> libbrotlienc.so.1 -> libbrotlienc.so.1.0.0 (= brotli_b200/libbrotlienc_b200.so, soname libbrotlienc.so.1)
#   <prefix>/lib/pkgconfig/libbrotlienc.pc, <prefix>/include/brotli/{encode,types,port,shared_dictionary}.h, <prefix>/include/brotli_b200.h
# usage: packaging/install_dropin.sh <prefix>
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"; PREFIX="$(mkdir -p "$1" && cd "$1" && pwd)"
mkdir -p "$PREFIX/lib/pkgconfig" "$PREFIX/include/brotli"
cp "$ROOT/brotli_b200/libbrotlienc_b200.so" "$PREFIX/lib/libbrotlienc.so.1.0.0"
ln -sf libbrotlienc.so.1.0.0 "$PREFIX/lib/libbrotlienc.so.1"; ln -sf libbrotlienc.so.1 "$PREFIX/lib/libbrotlienc.so"
cp "$ROOT"/include/brotli/*.h "$PREFIX/include/brotli/"; cp "$ROOT/include/brotli_b200.h" "$PREFIX/include/"
sed "s|@prefix@|$PREFIX|" "$ROOT/packaging/libbrotlienc.pc.in" > "$PREFIX/lib/pkgconfig/libbrotlienc.pc"
echo "installed under $PREFIX"
