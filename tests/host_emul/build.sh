#!/bin/sh
# Builds the test-only host harness (see emul.cpp).  Not part of the product.
set -e
cd "$(dirname "$0")"
g++ -O2 -std=c++17 -shared -fPIC -o libdgmk_emul.so emul.cpp -lm
